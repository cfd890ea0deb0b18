#ifndef REF_SHIM_BOOST_UTILITY
#define REF_SHIM_BOOST_UTILITY
#include <type_traits>
namespace boost {
template <class Cond, class T = void> struct disable_if : std::enable_if<!Cond::value, T> {};
template <class Cond, class T = void> struct enable_if : std::enable_if<Cond::value, T> {};
template <class T> struct is_pointer : std::is_pointer<T> {};
}

/* boost::math::sign (boost/math/special_functions/sign.hpp: -1, 0 or +1); h_signature.h:373 uses it without including it */
namespace boost { namespace math {
template <typename T> inline int sign(const T& z) { return (z == 0) ? 0 : ((z < 0) ? -1 : 1); }
} }
#endif
