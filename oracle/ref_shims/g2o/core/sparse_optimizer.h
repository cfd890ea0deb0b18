// g2o::SparseOptimizer + BlockSolver + OptimizationAlgorithmLevenberg + LinearSolverCSparse reduced to one class:
// a restatement of the g2o behaviour the reference relies on (upstream g2o 2018.3.25 /
// 2020.5.3 as packaged by ROS melodic / noetic - third-party code that is NOT part of /root/reference):
//   initializeOptimization  active vertices by id, hessian indices for non-fixed vertices, active edges by insertion
//   optimize(n)             Levenberg-Marquardt: tau 1e-5, nu doubling, 10 trials, [1/3, 2/3] clamp, +1e-3 scale,
//                           Terminate semantics; lambda / nu restart at every optimize() call
//   linear solver           banded Cholesky on the scalar system (any exact SPD solver returns the same step up to
//                           round-off; CSparse's failure semantics - x left equal to b - are kept)
// The arithmetic order mirrors oracle/teb_oracle.c on purpose: whatever differs between oracle/_ref and the
// restatement then comes from the REFERENCE-sourced code (edges, graph construction, autoResize, cost), which is what
// oracle/_ref exists to pin. TEST INFRASTRUCTURE.
#ifndef REF_SHIM_G2O_SPARSE_OPTIMIZER
#define REF_SHIM_G2O_SPARSE_OPTIMIZER
#include <g2o/core/base_vertex.h>
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>
namespace g2o {
struct G2OBatchStatistics {
  int iteration;
  double chi2;
  G2OBatchStatistics() : iteration(0), chi2(0) {}
};
typedef std::vector<G2OBatchStatistics> BatchStatisticsContainer;
class OptimizationAlgorithm {
 public:
  virtual ~OptimizationAlgorithm() {}
};
class SparseOptimizer {
 public:
  typedef std::vector<OptimizableGraph::Edge*> EdgeContainer;
  SparseOptimizer() : _alg(NULL), _verbose(false), _stats(false), _nextEdgeId(0), _N(0), _W(1), lm_trials(0), lm_rejected(0), last_terminated(false), chol_failed(false) {}
  ~SparseOptimizer() { clear(); delete _alg; }
  HyperGraph::VertexIDMap& vertices() { return _vertices; }
  HyperGraph::EdgeSet& edges() { return _edges; }
  const EdgeContainer& activeEdges() const { return _activeEdges; }
  bool addVertex(OptimizableGraph::Vertex* v) { _vertices[v->id()] = v; return true; }
  bool addEdge(OptimizableGraph::Edge* e) {
    e->setInternalId(_nextEdgeId++);
    _edges.insert(e);
    for (size_t i = 0; i < e->numVertices(); ++i) e->vertexAt(i)->edges().insert(e);
    return true;
  }
  void setAlgorithm(OptimizationAlgorithm* a) { delete _alg; _alg = a; }
  void initMultiThreading() {}
  void setVerbose(bool v) { _verbose = v; }
  void setComputeBatchStatistics(bool s) { _stats = s; }
  const BatchStatisticsContainer& batchStatistics() const { return _batch; }
  void computeInitialGuess() {} /* no TEB edge implements initialEstimate */
  void clear() { /* HyperGraph::clear(): deletes what is still registered */
    for (HyperGraph::VertexIDMap::iterator it = _vertices.begin(); it != _vertices.end(); ++it) delete it->second;
    for (HyperGraph::EdgeSet::iterator it = _edges.begin(); it != _edges.end(); ++it) delete *it;
    _vertices.clear(); _edges.clear(); _activeEdges.clear(); _activeVertices.clear();
  }
  bool initializeOptimization(int = 0) {
    _activeVertices.clear();
    for (HyperGraph::VertexIDMap::iterator it = _vertices.begin(); it != _vertices.end(); ++it)
      _activeVertices.push_back(static_cast<OptimizableGraph::Vertex*>(it->second)); /* std::map: ascending id */
    _activeEdges.clear();
    for (HyperGraph::EdgeSet::iterator it = _edges.begin(); it != _edges.end(); ++it) {
      OptimizableGraph::Edge* e = static_cast<OptimizableGraph::Edge*>(*it);
      bool allFixed = true;
      for (size_t i = 0; i < e->numVertices(); ++i) allFixed = allFixed && e->vertexAt(i)->fixed();
      if (!allFixed) _activeEdges.push_back(e);
    }
    std::sort(_activeEdges.begin(), _activeEdges.end(),
              [](const OptimizableGraph::Edge* a, const OptimizableGraph::Edge* b) { return a->internalId() < b->internalId(); });
    _N = 0;
    for (size_t k = 0; k < _activeVertices.size(); ++k) {
      OptimizableGraph::Vertex* v = _activeVertices[k];
      if (v->fixed()) { v->setHessianIndex(-1); continue; }
      v->setHessianIndex(_N);
      _N += v->dimension();
    }
    int hbw = 0; /* half bandwidth of the scalar system */
    for (size_t k = 0; k < _activeEdges.size(); ++k) {
      int lo = 1 << 30, hi = -1;
      OptimizableGraph::Edge* e = _activeEdges[k];
      for (size_t i = 0; i < e->numVertices(); ++i) {
        OptimizableGraph::Vertex* v = e->vertexAt(i);
        if (v->fixed()) continue;
        lo = std::min(lo, v->hessianIndex());
        hi = std::max(hi, v->hessianIndex() + v->dimension() - 1);
      }
      if (hi >= lo) hbw = std::max(hbw, hi - lo);
    }
    _W = hbw + 1;
    _Hb.assign((size_t)_N * _W, 0.0); _L.assign((size_t)_N * _W, 0.0); _b.assign(_N, 0.0); _x.assign(_N, 0.0);
    return true;
  }
  int optimize(int iterations, bool = false) {
    _batch.clear();
    double lambda = 0, ni = 2;
    int ok = 1, cj = 0;
    for (int i = 0; i < iterations && ok; ++i) {
      ok = solveLM(i, lambda, ni);
      if (_stats) { G2OBatchStatistics s; s.iteration = i; s.chi2 = computeActiveErrors(); _batch.push_back(s); }
      ++cj;
    }
    last_terminated = !ok;
    return cj;
  }
  double computeActiveErrors() {
    double chi2 = 0;
    for (size_t k = 0; k < _activeEdges.size(); ++k) {
      OptimizableGraph::Edge* e = _activeEdges[k];
      e->computeError();
      const int D = e->dimension();
      const double* er = e->errorData();
      const double* W = e->informationData();
      for (int d = 0; d < D; ++d) chi2 += er[d] * W[d + D * d] * er[d]; /* the reference only sets diagonal information */
    }
    return chi2;
  }
  /* test hooks */
  int systemSize() const { return _N; }
  int halfBandwidth() const { return _W - 1; }
  const std::vector<double>& bandedH() const { return _Hb; }
  const std::vector<double>& rhs() const { return _b; }
  void buildSystemOnly() { computeActiveErrors(); buildSystem(); }
  long lm_trials, lm_rejected;
  bool last_terminated, chol_failed;

 private:
  void Hadd(int r, int q, double v) { if (r < q) std::swap(r, q); _Hb[(size_t)r * _W + (r - q)] += v; }
  void buildSystem() { /* BlockSolver::buildSystem: linearizeOplus + constructQuadraticForm per edge (App. A.3) */
    std::fill(_Hb.begin(), _Hb.end(), 0.0);
    std::fill(_b.begin(), _b.end(), 0.0);
    for (size_t k = 0; k < _activeEdges.size(); ++k) {
      OptimizableGraph::Edge* e = _activeEdges[k];
      e->linearizeOplus();
      const int D = e->dimension();
      const double* er = e->errorData();
      const double* W = e->informationData();
      double omega_r[8];
      for (int d = 0; d < D; ++d) omega_r[d] = -W[d + D * d] * er[d];
      const size_t nv = e->numVertices();
      for (size_t i = 0; i < nv; ++i) {
        OptimizableGraph::Vertex* vi = e->vertexAt(i);
        if (vi->fixed()) continue;
        const int di = vi->dimension(), hi = vi->hessianIndex();
        const double* A = e->jacobianData(i); /* column major D x di: A[d + D a] */
        for (int a = 0; a < di; ++a) {
          double s = 0;
          for (int d = 0; d < D; ++d) s += A[d + D * a] * omega_r[d];
          _b[hi + a] += s;
        }
        for (int a = 0; a < di; ++a)
          for (int bq = 0; bq <= a; ++bq) {
            double s = 0;
            for (int d = 0; d < D; ++d) s += A[d + D * a] * W[d + D * d] * A[d + D * bq];
            Hadd(hi + a, hi + bq, s);
          }
        for (size_t j = i + 1; j < nv; ++j) {
          OptimizableGraph::Vertex* vj = e->vertexAt(j);
          if (vj->fixed()) continue;
          const int dj = vj->dimension(), hj = vj->hessianIndex();
          const double* Bm = e->jacobianData(j);
          for (int a = 0; a < di; ++a)
            for (int bq = 0; bq < dj; ++bq) {
              double s = 0;
              for (int d = 0; d < D; ++d) s += A[d + D * a] * W[d + D * d] * Bm[d + D * bq];
              Hadd(hi + a, hj + bq, s);
            }
        }
      }
    }
  }
  bool solveBanded(double lambda) { /* banded Cholesky LL^T, lower storage */
    const int N = _N, W = _W, HBW = _W - 1;
    std::vector<double>& L = _L;
    L = _Hb;
    for (int r = 0; r < N; ++r) L[(size_t)r * W] += lambda;
    for (int j = 0; j < N; ++j) {
      const int kmax = j < HBW ? j : HBW;
      for (int k = kmax; k >= 0; --k) {
        const int col = j - k;
        double s = L[(size_t)j * W + k];
        int m0 = j - HBW; if (m0 < 0) m0 = 0;
        for (int m = m0; m < col; ++m) s -= L[(size_t)j * W + (j - m)] * L[(size_t)col * W + (col - m)];
        if (k == 0) {
          if (!(s > 0) || !std::isfinite(s)) return false;
          L[(size_t)j * W] = std::sqrt(s);
        } else {
          L[(size_t)j * W + k] = s / L[(size_t)col * W];
        }
      }
    }
    for (int j = 0; j < N; ++j) {
      double s = _b[j];
      int m0 = j - HBW; if (m0 < 0) m0 = 0;
      for (int m = m0; m < j; ++m) s -= L[(size_t)j * W + (j - m)] * _x[m];
      _x[j] = s / L[(size_t)j * W];
    }
    for (int j = N - 1; j >= 0; --j) {
      double s = _x[j];
      int m1 = j + HBW; if (m1 > N - 1) m1 = N - 1;
      for (int m = j + 1; m <= m1; ++m) s -= L[(size_t)m * W + (m - j)] * _x[m];
      _x[j] = s / L[(size_t)j * W];
    }
    return true;
  }
  void push() { for (size_t k = 0; k < _activeVertices.size(); ++k) _activeVertices[k]->push(); }
  void pop() { for (size_t k = 0; k < _activeVertices.size(); ++k) _activeVertices[k]->pop(); }
  void discardTop() { for (size_t k = 0; k < _activeVertices.size(); ++k) _activeVertices[k]->discardTop(); }
  void update() { /* SparseOptimizer::update: every non-fixed vertex oplus its slice of x */
    for (size_t k = 0; k < _activeVertices.size(); ++k) {
      OptimizableGraph::Vertex* v = _activeVertices[k];
      if (v->hessianIndex() >= 0) v->oplus(&_x[v->hessianIndex()]);
    }
  }
  int solveLM(int iteration, double& lambda, double& ni) { /* OptimizationAlgorithmLevenberg::solve (App. A.4) */
    double currentChi = computeActiveErrors();
    double tempChi = currentChi;
    buildSystem();
    if (iteration == 0) {
      double maxDiagonal = 0;
      for (int r = 0; r < _N; ++r) maxDiagonal = std::max(maxDiagonal, std::fabs(_Hb[(size_t)r * _W]));
      lambda = 1e-5 * maxDiagonal;
      ni = 2;
    }
    double rho = 0;
    int qmax = 0;
    do {
      push();
      const bool ok2 = solveBanded(lambda);
      if (!ok2) { chol_failed = true; _x = _b; } /* CSparse leaves x = b when the factorisation fails */
      ++lm_trials;
      update();
      tempChi = computeActiveErrors();
      if (!ok2) tempChi = DBL_MAX;
      rho = (currentChi - tempChi);
      double scale = 0;
      for (int j = 0; j < _N; ++j) scale += _x[j] * (lambda * _x[j] + _b[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = alpha < (2. / 3.) ? alpha : (2. / 3.);
        const double scaleFactor = (1. / 3.) > alpha ? (1. / 3.) : alpha;
        lambda *= scaleFactor;
        ni = 2;
        currentChi = tempChi;
        discardTop();
      } else {
        lambda *= ni;
        ni *= 2;
        pop();
        ++lm_rejected;
        if (!std::isfinite(lambda)) break;
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    if (qmax == 10 || rho == 0 || !std::isfinite(lambda)) return 0;
    return 1;
  }
  HyperGraph::VertexIDMap _vertices;
  HyperGraph::EdgeSet _edges;
  EdgeContainer _activeEdges;
  std::vector<OptimizableGraph::Vertex*> _activeVertices;
  OptimizationAlgorithm* _alg;
  bool _verbose, _stats;
  long long _nextEdgeId;
  int _N, _W;
  std::vector<double> _Hb, _L, _b, _x;
  BatchStatisticsContainer _batch;
};
/* the solver-configuration types optimal_planner.{h,cpp} names (optimal_planner.cpp:161-179) */
template <int P, int L> struct BlockSolverTraits { typedef int PoseMatrixType; };
template <typename MatrixType> class LinearSolverCSparse { public: void setBlockOrdering(bool) {} };
template <typename MatrixType> class LinearSolverCholmod { public: void setBlockOrdering(bool) {} };
template <typename Traits> class BlockSolver {
 public:
  typedef typename Traits::PoseMatrixType PoseMatrixType;
  template <typename LS> explicit BlockSolver(std::unique_ptr<LS>) {}
};
class OptimizationAlgorithmLevenberg : public OptimizationAlgorithm {
 public:
  template <typename BS> explicit OptimizationAlgorithmLevenberg(std::unique_ptr<BS>) {}
};
class OptimizationAlgorithmGaussNewton : public OptimizationAlgorithm {
 public:
  template <typename BS> explicit OptimizationAlgorithmGaussNewton(std::unique_ptr<BS>) {}
};
class AbstractHyperGraphElementCreator { public: virtual ~AbstractHyperGraphElementCreator() {} };
template <typename T> class HyperGraphElementCreator : public AbstractHyperGraphElementCreator {};
class Factory {
 public:
  static Factory* instance() { static Factory f; return &f; }
  void registerType(const std::string&, AbstractHyperGraphElementCreator* c) { delete c; }
  static void destroy() {}
};
}  // namespace g2o
#endif
