#ifndef REF_SHIM_G2O_UNARY
#define REF_SHIM_G2O_UNARY
#include <g2o/core/base_vertex.h>
namespace g2o {
template <int D, typename E, typename VertexXi>
class BaseUnaryEdge : public BaseEdge<D, E> {
 public:
  typedef Eigen::Matrix<double, D, VertexXi::Dimension> JacobianXiOplusType;
  using typename BaseEdge<D, E>::ErrorVector;
  BaseUnaryEdge() { this->_vertices.resize(1, NULL); }
  virtual const double* jacobianData(size_t i) const { return this->_numeric ? this->_jnum[i].data() : _jacobianOplusXi.data(); }
 protected:
  JacobianXiOplusType _jacobianOplusXi; /* filled by the analytic overrides of the reference */
};
}
#endif
