#ifndef REF_SHIM_BOOST_ONCE
#define REF_SHIM_BOOST_ONCE
#include <mutex>
namespace boost {
typedef std::once_flag once_flag;
#define BOOST_ONCE_INIT {}
template <typename F> inline void call_once(F f, once_flag& flag) { std::call_once(flag, f); }
}
#endif
