// Stand-in for boost::random::mt19937 and boost::random::uniform_real_distribution<double> (TEST INFRASTRUCTURE).
// boost is not part of /root/reference: the engine is the standard mt19937 (same algorithm, default seed 5489); the
// distribution restates boost/random/uniform_real_distribution.hpp (detail::generate_uniform_real for an integer engine):
// ONE 32-bit draw per sample, value = (draw - min) / (max - min + 1) * (b - a) + a, redrawn when it reaches b. That is
// NOT what libstdc++'s std::uniform_real_distribution does (two draws, generate_canonical), hence the restatement.
#ifndef REF_SHIM_BOOST_RANDOM
#define REF_SHIM_BOOST_RANDOM
#include <random>
namespace boost { namespace random {
typedef std::mt19937 mt19937;
template <class T = double> class uniform_real_distribution {
 public:
  uniform_real_distribution(T a = 0, T b = 1) : a_(a), b_(b) {}
  template <class Engine> T operator()(Engine& eng) const {
    for (;;) {
      const T numerator = static_cast<T>(eng() - (Engine::min)());
      const T divisor = static_cast<T>((Engine::max)() - (Engine::min)()) + 1;
      const T result = numerator / divisor * (b_ - a_) + a_;
      if (result < b_) return result;
    }
  }
 private:
  T a_, b_;
};
} using random::mt19937; }
#endif
