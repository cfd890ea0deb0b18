#ifndef REF_SHIM_BOOST_OPTIONAL
#define REF_SHIM_BOOST_OPTIONAL
#include <iterator>
namespace boost {
struct none_t {};
static const none_t none = none_t();
template <typename T> class optional {
 public:
  optional() : has_(false), v_() {}
  optional(none_t) : has_(false), v_() {}
  optional(const T& v) : has_(true), v_(v) {}
  explicit operator bool() const { return has_; }
  const T& operator*() const { return v_; }
  T& operator*() { return v_; }
  const T* operator->() const { return &v_; }
  const T& get() const { return v_; }
 private:
  bool has_;
  T v_;
};
template <typename T> class optional<T&> {
 public:
  optional() : p_(0) {}
  optional(none_t) : p_(0) {}
  optional(T& v) : p_(&v) {}
  explicit operator bool() const { return p_ != 0; }
  T& operator*() const { return *p_; }
  T* operator->() const { return p_; }
  T& get() const { return *p_; }
 private:
  T* p_;
};
template <typename T> inline bool operator==(const optional<T>& o, none_t) { return !o; }
template <typename T> inline bool operator!=(const optional<T>& o, none_t) { return bool(o); }
template <typename It> inline It prior(It it) { return std::prev(it); }
template <typename It> inline It next(It it) { return std::next(it); }
}
#endif
