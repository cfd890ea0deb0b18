#ifndef REF_SHIM_G2O_MULTI
#define REF_SHIM_G2O_MULTI
#include <g2o/core/base_vertex.h>
namespace g2o {
template <int D, typename E>
class BaseMultiEdge : public BaseEdge<D, E> {
 public:
  using typename BaseEdge<D, E>::ErrorVector;
};
}
#endif
