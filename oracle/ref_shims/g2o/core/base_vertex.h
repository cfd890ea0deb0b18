// g2o vertex / edge base classes reduced to what the reference's g2o_types/*.h use (base_teb_edges.h:47-51):
// _estimate / oplus / push / pop / fixed for vertices; _error, _measurement, _vertices, _information,
// _jacobianOplusXi / Xj, computeError(), linearizeOplus() with g2o's numeric default (central differences,
// delta = 1e-9, fixed vertices skipped) for edges. TEST INFRASTRUCTURE.
#ifndef REF_SHIM_G2O_BASE
#define REF_SHIM_G2O_BASE
#include <Eigen/Core>
#include <cstring>
#include <iostream>
#include <map>
#include <set>
#include <stack>
#include <vector>
namespace g2o {
class HyperGraph {
 public:
  class Edge;
  typedef std::set<Edge*> EdgeSet;
  class Vertex {
   public:
    virtual ~Vertex() {}
    EdgeSet& edges() { return _edges; }
    const EdgeSet& edges() const { return _edges; }
   protected:
    EdgeSet _edges;
  };
  class Edge {
   public:
    Edge() : _internalId(-1) {}
    virtual ~Edge() {}
    long long internalId() const { return _internalId; }
    void setInternalId(long long i) { _internalId = i; }
   protected:
    long long _internalId;
  };
  typedef std::map<int, Vertex*> VertexIDMap;
};
class OptimizableGraph {
 public:
  class Vertex : public HyperGraph::Vertex {
   public:
    Vertex() : _fixed(false), _id(-1), _hessianIndex(-1) {}
    bool fixed() const { return _fixed; }
    void setFixed(bool f) { _fixed = f; }
    int id() const { return _id; }
    void setId(int i) { _id = i; }
    int hessianIndex() const { return _hessianIndex; }
    void setHessianIndex(int i) { _hessianIndex = i; }
    virtual int dimension() const = 0;
    void oplus(const double* v) { oplusImpl(v); }
    virtual void push() = 0;
    virtual void pop() = 0;
    virtual void discardTop() = 0;
    virtual void setToOriginImpl() = 0;
    virtual void oplusImpl(const double* update) = 0;
    virtual bool read(std::istream& is) = 0;
    virtual bool write(std::ostream& os) const = 0;
   protected:
    bool _fixed;
    int _id, _hessianIndex;
  };
  class Edge : public HyperGraph::Edge {
   public:
    virtual void computeError() = 0;
    virtual void linearizeOplus() = 0;
    virtual int dimension() const = 0;
    virtual const double* errorData() const = 0;
    virtual const double* informationData() const = 0; /* D x D column major */
    virtual size_t numVertices() const = 0;
    virtual Vertex* vertexAt(size_t i) const = 0;
    virtual const double* jacobianData(size_t i) const = 0; /* D x dim_i, column major, valid after linearizeOplus() */
    virtual bool read(std::istream& is) = 0;
    virtual bool write(std::ostream& os) const = 0;
    double chi2() const {
      const int D = dimension();
      const double* e = errorData();
      const double* W = informationData();
      double s = 0;
      for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) s += e[i] * W[i + D * j] * e[j];
      return s;
    }
  };
};
template <int D, typename T>
class BaseVertex : public OptimizableGraph::Vertex {
 public:
  static const int Dimension = D;
  typedef T EstimateType;
  const T& estimate() const { return _estimate; }
  void setEstimate(const T& e) { _estimate = e; }
  virtual int dimension() const { return D; }
  virtual void push() { _backup.push(_estimate); }
  virtual void pop() { _estimate = _backup.top(); _backup.pop(); }
  virtual void discardTop() { _backup.pop(); }
 protected:
  T _estimate;
  std::stack<T> _backup;
};
/* numeric linearisation shared by the three edge arities (g2o base_*_edge.hpp: delta = 1e-9, scalar = 1 / (2 delta)) */
template <int D>
struct NumericLinearizer {
  template <typename EdgeT>
  static void run(EdgeT* edge, Eigen::Matrix<double, D, 1>& error, std::vector<std::vector<double> >& jac) {
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    const size_t nv = edge->numVertices();
    jac.resize(nv);
    const Eigen::Matrix<double, D, 1> backup = error;
    for (size_t i = 0; i < nv; ++i) {
      OptimizableGraph::Vertex* v = edge->vertexAt(i);
      const int dim = v->dimension();
      jac[i].assign((size_t)D * dim, 0.0);
      if (v->fixed()) continue;
      double add[8];
      for (int d = 0; d < dim; ++d) {
        std::memset(add, 0, sizeof(add));
        v->push();
        add[d] = delta;
        v->oplus(add);
        edge->computeError();
        const Eigen::Matrix<double, D, 1> e1 = error;
        v->pop();
        v->push();
        add[d] = -delta;
        v->oplus(add);
        edge->computeError();
        const Eigen::Matrix<double, D, 1> e2 = error;
        v->pop();
        for (int r = 0; r < D; ++r) jac[i][r + (size_t)D * d] = scalar * (e1[r] - e2[r]);
      }
    }
    error = backup;
  }
};
template <int D, typename E>
class BaseEdge : public OptimizableGraph::Edge {
 public:
  typedef Eigen::Matrix<double, D, 1> ErrorVector;
  typedef Eigen::Matrix<double, D, D> InformationType;
  typedef E Measurement;
  BaseEdge() { _information.setIdentity(); }
  virtual int dimension() const { return D; }
  const ErrorVector& error() const { return _error; }
  ErrorVector& error() { return _error; }
  const InformationType& information() const { return _information; }
  void setInformation(const InformationType& i) { _information = i; }
  const E& measurement() const { return _measurement; }
  void setMeasurement(const E& m) { _measurement = m; }
  virtual const double* errorData() const { return _error.data(); }
  virtual const double* informationData() const { return _information.data(); }
  virtual size_t numVertices() const { return _vertices.size(); }
  virtual OptimizableGraph::Vertex* vertexAt(size_t i) const { return static_cast<OptimizableGraph::Vertex*>(_vertices[i]); }
  void setVertex(size_t i, HyperGraph::Vertex* v) { _vertices[i] = v; }
  const std::vector<HyperGraph::Vertex*>& vertices() const { return _vertices; }
  virtual void resize(size_t n) { _vertices.resize(n, NULL); }
  virtual void linearizeOplus() { NumericLinearizer<D>::run(this, _error, _jnum); _numeric = true; }
  virtual const double* jacobianData(size_t i) const { return _jnum[i].data(); }
 protected:
  ErrorVector _error;
  InformationType _information;
  E _measurement;
  std::vector<HyperGraph::Vertex*> _vertices;
  std::vector<std::vector<double> > _jnum;
  bool _numeric = false;
};
}  // namespace g2o
#endif
