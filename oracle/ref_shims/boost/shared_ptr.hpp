#ifndef REF_SHIM_BOOST_SHARED_PTR
#define REF_SHIM_BOOST_SHARED_PTR
#include <memory>
namespace boost {
using std::shared_ptr;
using std::make_shared;
using std::dynamic_pointer_cast;
using std::static_pointer_cast;
using std::const_pointer_cast;
}
#endif
