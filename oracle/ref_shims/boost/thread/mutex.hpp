#ifndef REF_SHIM_BOOST_MUTEX
#define REF_SHIM_BOOST_MUTEX
#include <mutex>
namespace boost { typedef std::mutex mutex; }
#endif
