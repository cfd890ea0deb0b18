// Stand-in for the part of boost::adjacency_list<listS, vecS, directedS, VertexProperty, no_property> that
// graph_search.{h,cpp} uses (TEST INFRASTRUCTURE: lets oracle/_ref compile the reference's graph search unmodified).
// vecS vertices: descriptors are indices, vertex iterators count them (random access: the reference writes `end - 1`);
// listS out-edges: adjacent_vertices() walks the targets in insertion order - what boost does for listS.
#ifndef REF_SHIM_BOOST_GRAPH_ADJACENCY_LIST
#define REF_SHIM_BOOST_GRAPH_ADJACENCY_LIST
#include <cstddef>
#include <iterator>
#include <list>
#include <tuple>
#include <utility>
#include <vector>
#include <boost/bind.hpp>   /* reaches graph_search.cpp transitively in the real tree */
#include <boost/thread.hpp>
namespace boost {
struct listS {}; struct vecS {}; struct directedS {}; struct no_property {};
template <class OutEdgeList, class VertexList, class Directed, class VertexProperty, class EdgeProperty = no_property>
class adjacency_list {
 public:
  typedef std::size_t vertex_descriptor;
  struct edge_descriptor { std::size_t source, target; };
  class vertex_iterator {
   public:
    typedef std::random_access_iterator_tag iterator_category;
    typedef std::size_t value_type; typedef std::ptrdiff_t difference_type; typedef const std::size_t* pointer; typedef std::size_t reference;
    vertex_iterator(std::size_t i = 0) : i_(i) {}
    std::size_t operator*() const { return i_; }
    vertex_iterator& operator++() { ++i_; return *this; }
    vertex_iterator operator++(int) { vertex_iterator t(*this); ++i_; return t; }
    vertex_iterator& operator--() { --i_; return *this; }
    vertex_iterator& operator+=(difference_type d) { i_ += d; return *this; }
    vertex_iterator& operator-=(difference_type d) { i_ -= d; return *this; }
    vertex_iterator operator-(difference_type d) const { return vertex_iterator(i_ - d); }
    vertex_iterator operator+(difference_type d) const { return vertex_iterator(i_ + d); }
    difference_type operator-(const vertex_iterator& o) const { return (difference_type)i_ - (difference_type)o.i_; }
    bool operator==(const vertex_iterator& o) const { return i_ == o.i_; }
    bool operator!=(const vertex_iterator& o) const { return i_ != o.i_; }
   private:
    std::size_t i_;
  };
  typedef typename std::list<std::size_t>::const_iterator adjacency_iterator;
  typedef adjacency_iterator out_edge_iterator;
  typedef vertex_iterator edge_iterator; /* only named by a typedef of the reference */
  VertexProperty& operator[](vertex_descriptor v) { return props_[v]; }
  const VertexProperty& operator[](vertex_descriptor v) const { return props_[v]; }
  void clear() { props_.clear(); adj_.clear(); }
  std::vector<VertexProperty> props_;
  std::vector<std::list<std::size_t> > adj_;
};
template <class G> struct graph_traits {
  typedef typename G::vertex_descriptor vertex_descriptor;
  typedef typename G::edge_descriptor edge_descriptor;
  typedef typename G::vertex_iterator vertex_iterator;
  typedef typename G::edge_iterator edge_iterator;
  typedef typename G::adjacency_iterator adjacency_iterator;
};
template <class A, class B, class C, class D, class E>
inline std::size_t add_vertex(adjacency_list<A, B, C, D, E>& g) { g.props_.push_back(D()); g.adj_.push_back(std::list<std::size_t>()); return g.props_.size() - 1; }
template <class A, class B, class C, class D, class E>
inline std::pair<typename adjacency_list<A, B, C, D, E>::edge_descriptor, bool> add_edge(std::size_t u, std::size_t v, adjacency_list<A, B, C, D, E>& g) {
  g.adj_[u].push_back(v);
  typename adjacency_list<A, B, C, D, E>::edge_descriptor e; e.source = u; e.target = v;
  return std::make_pair(e, true);
}
template <class A, class B, class C, class D, class E>
inline std::size_t num_vertices(const adjacency_list<A, B, C, D, E>& g) { return g.props_.size(); }
template <class A, class B, class C, class D, class E>
inline std::pair<typename adjacency_list<A, B, C, D, E>::vertex_iterator, typename adjacency_list<A, B, C, D, E>::vertex_iterator>
vertices(const adjacency_list<A, B, C, D, E>& g) {
  typedef typename adjacency_list<A, B, C, D, E>::vertex_iterator It;
  return std::make_pair(It(0), It(g.props_.size()));
}
template <class A, class B, class C, class D, class E>
inline std::pair<typename adjacency_list<A, B, C, D, E>::adjacency_iterator, typename adjacency_list<A, B, C, D, E>::adjacency_iterator>
adjacent_vertices(std::size_t v, const adjacency_list<A, B, C, D, E>& g) { return std::make_pair(g.adj_[v].begin(), g.adj_[v].end()); }
template <class T1, class T2> inline std::tuple<T1&, T2&> tie(T1& a, T2& b) { return std::tuple<T1&, T2&>(a, b); }
}  // namespace boost
#endif
