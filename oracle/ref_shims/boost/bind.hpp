// boost::bind / boost::cref / _1 .. _3 on top of <functional> (TEST INFRASTRUCTURE)
#ifndef REF_SHIM_BOOST_BIND
#define REF_SHIM_BOOST_BIND
#include <functional>
namespace boost { using std::bind; using std::cref; using std::ref; }
namespace { const decltype(std::placeholders::_1)& _1 = std::placeholders::_1; const decltype(std::placeholders::_2)& _2 = std::placeholders::_2;
            const decltype(std::placeholders::_3)& _3 = std::placeholders::_3; }
#endif
