#ifndef REF_SHIM_BOOST_UTILITY
#define REF_SHIM_BOOST_UTILITY
#include <type_traits>
namespace boost {
template <class Cond, class T = void> struct disable_if : std::enable_if<!Cond::value, T> {};
template <class Cond, class T = void> struct enable_if : std::enable_if<Cond::value, T> {};
template <class T> struct is_pointer : std::is_pointer<T> {};
}
#endif
