#ifndef REF_SHIM_G2O_BINARY
#define REF_SHIM_G2O_BINARY
#include <g2o/core/base_vertex.h>
namespace g2o {
template <int D, typename E, typename VertexXi, typename VertexXj>
class BaseBinaryEdge : public BaseEdge<D, E> {
 public:
  typedef Eigen::Matrix<double, D, VertexXi::Dimension> JacobianXiOplusType;
  typedef Eigen::Matrix<double, D, VertexXj::Dimension> JacobianXjOplusType;
  using typename BaseEdge<D, E>::ErrorVector;
  BaseBinaryEdge() { this->_vertices.resize(2, NULL); }
  virtual const double* jacobianData(size_t i) const {
    if (this->_numeric) return this->_jnum[i].data();
    return i == 0 ? _jacobianOplusXi.data() : _jacobianOplusXj.data();
  }
 protected:
  JacobianXiOplusType _jacobianOplusXi;
  JacobianXjOplusType _jacobianOplusXj;
};
}
#endif
