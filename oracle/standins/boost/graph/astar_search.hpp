// Boost.Graph API stand-in (oracle/standins/README.md): astar_search restated from the library's published
// algorithm: the initialising overload, breadth_first_visit driven by a 4-ary indirect heap on
// the f-cost, the astar_bfs_visitor events, relax() with its undirected reverse branch, closed_plus.
#pragma once
#include <functional>
#include <stdexcept>
#include <vector>
#include "boost/graph/adjacency_list.hpp"
namespace boost {
struct negative_edge : public std::invalid_argument {
  negative_edge() : std::invalid_argument("The graph may not contain an edge with negative weight.") {}
};
template <class T> struct closed_plus {
  const T inf;
  closed_plus() : inf(std::numeric_limits<T>::max()) {}
  explicit closed_plus(T i) : inf(i) {}
  T operator()(const T& a, const T& b) const {
    if (a == inf) return inf;
    if (b == inf) return inf;
    return a + b;
  }
};
// As in Boost: operator() is NOT virtual and returns zero; astar_search takes the heuristic BY VALUE.
template <class Graph, class CostType> class astar_heuristic {
public:
  typedef typename graph_traits<Graph>::vertex_descriptor Vertex;
  typedef Vertex argument_type;
  typedef CostType result_type;
  astar_heuristic() {}
  CostType operator()(Vertex) { return static_cast<CostType>(0); }
};
template <class Visitors = int> class default_astar_visitor {
public:
  template <class V, class G> void initialize_vertex(V, const G&) {}
  template <class V, class G> void discover_vertex(V, const G&) {}
  template <class V, class G> void examine_vertex(V, const G&) {}
  template <class E, class G> void examine_edge(E, const G&) {}
  template <class E, class G> void edge_relaxed(E, const G&) {}
  template <class E, class G> void edge_not_relaxed(E, const G&) {}
  template <class E, class G> void black_target(E, const G&) {}
  template <class V, class G> void finish_vertex(V, const G&) {}
};
namespace standin_detail {
// d_ary_heap_indirect<Vertex, 4, IndexInHeap, CostMap, Compare>: values are vertices, keys live in the cost
// map; push and update (decrease-key) sift up, pop moves the last element to the root and sifts down; among
// equal children the first smallest wins.
template <class CostMap, class Compare> class heap4 {
public:
  heap4(CostMap cost, Compare cmp, std::size_t n) : cost_(cost), cmp_(cmp), pos_(n, (std::size_t)-1) {}
  bool empty() const { return data_.empty(); }
  std::size_t top() const { return data_[0]; }
  void push(std::size_t v) {
    if (v >= pos_.size()) pos_.resize(v + 1, (std::size_t)-1);
    std::size_t i = data_.size();
    data_.push_back(v);
    pos_[v] = i;
    up(i);
  }
  void update(std::size_t v) { up(pos_[v]); }
  void pop() {
    pos_[data_[0]] = (std::size_t)-1;
    if (data_.size() != 1) {
      data_[0] = data_.back();
      pos_[data_[0]] = 0;
      data_.pop_back();
      down();
    } else {
      data_.pop_back();
    }
  }
private:
  void up(std::size_t index) {
    std::size_t orig = index, levels = 0;
    if (index == 0) return;
    std::size_t moving = data_[index];
    auto moving_dist = get(cost_, moving);
    for (;;) {
      if (index == 0) break;
      std::size_t parent = (index - 1) / 4;
      if (cmp_(moving_dist, get(cost_, data_[parent]))) {
        ++levels;
        index = parent;
        continue;
      }
      break;
    }
    index = orig;
    for (std::size_t i = 0; i < levels; ++i) {
      std::size_t parent = (index - 1) / 4;
      data_[index] = data_[parent];
      pos_[data_[index]] = index;
      index = parent;
    }
    data_[index] = moving;
    pos_[moving] = index;
  }
  void down() {
    if (data_.empty()) return;
    std::size_t index = 0;
    std::size_t moving = data_[0];
    auto moving_dist = get(cost_, moving);
    std::size_t heap_size = data_.size();
    for (;;) {
      std::size_t first_child = 4 * index + 1;
      if (first_child >= heap_size) break;
      std::size_t smallest = 0;
      auto smallest_dist = get(cost_, data_[first_child]);
      std::size_t nchildren = first_child + 4 <= heap_size ? 4 : heap_size - first_child;
      for (std::size_t i = 1; i < nchildren; ++i) {
        auto d = get(cost_, data_[first_child + i]);
        if (cmp_(d, smallest_dist)) {
          smallest = i;
          smallest_dist = d;
        }
      }
      if (cmp_(smallest_dist, moving_dist)) {
        std::size_t child = first_child + smallest;
        // swap_heap_elements(child, index)
        std::size_t a = data_[child], b = data_[index];
        data_[child] = b;
        data_[index] = a;
        pos_[a] = index;
        pos_[b] = child;
        index = child;
        continue;
      }
      break;
    }
  }
  CostMap cost_;
  Compare cmp_;
  std::vector<std::size_t> data_, pos_;
};
template <class Graph, class Edge, class WeightMap, class PredMap, class DistMap, class Combine, class Compare>
inline bool relax(const Edge& e, const Graph& g, const WeightMap& w, PredMap& p, DistMap& d, const Combine& combine,
                  const Compare& compare) {
  const std::size_t u = source(e, g), v = target(e, g);
  const double d_u = get(d, u);
  const double d_v = get(d, v);
  const double w_e = get(w, e);
  if (compare(combine(d_u, w_e), d_v)) {
    put(d, v, combine(d_u, w_e));
    if (compare(get(d, v), d_v)) {
      put(p, v, u);
      return true;
    }
    return false;
  } else if (compare(combine(d_v, w_e), d_u)) {  // undirected graphs only (the only kind here)
    put(d, u, combine(d_v, w_e));
    if (compare(get(d, u), d_u)) {
      put(p, u, v);
      return true;
    }
    return false;
  }
  return false;
}
}  // namespace standin_detail

template <class Graph, class Heuristic, class Visitor, class PredMap, class CostMap, class DistMap, class WeightMap,
          class IndexMap, class ColorMap, class Compare, class Combine, class CostInf, class CostZero>
inline void astar_search(const Graph& g, typename graph_traits<Graph>::vertex_descriptor s, Heuristic h, Visitor vis,
                         PredMap predecessor, CostMap cost, DistMap distance, WeightMap weight, IndexMap /*index_map*/,
                         ColorMap color, Compare compare, Combine combine, CostInf inf, CostZero zero) {
  typedef typename graph_traits<Graph>::vertex_iterator VIter;
  VIter ui, ui_end;
  for (boost::tie(ui, ui_end) = vertices(g); ui != ui_end; ++ui) {
    put(color, *ui, white_color);
    put(distance, *ui, inf);
    put(cost, *ui, inf);
    put(predecessor, *ui, *ui);
    vis.initialize_vertex(*ui, g);
  }
  put(distance, s, zero);
  put(cost, s, h(s));
  standin_detail::heap4<CostMap, Compare> Q(cost, compare, num_vertices(g));
  // breadth_first_visit
  put(color, s, gray_color);
  vis.discover_vertex(s, g);
  Q.push(s);
  while (!Q.empty()) {
    const std::size_t u = Q.top();
    Q.pop();
    vis.examine_vertex(u, g);
    typename graph_traits<Graph>::out_edge_iterator ei, ei_end;
    for (boost::tie(ei, ei_end) = out_edges(u, g); ei != ei_end; ++ei) {
      const auto e = *ei;
      const std::size_t v = target(e, g);
      // astar_bfs_visitor::examine_edge
      if (compare(get(weight, e), zero)) throw negative_edge();
      vis.examine_edge(e, g);
      const default_color_type v_color = get(color, v);
      if (v_color == white_color) {  // tree_edge
        const bool decreased = standin_detail::relax(e, g, weight, predecessor, distance, combine, compare);
        if (decreased) {
          vis.edge_relaxed(e, g);
          put(cost, v, combine(get(distance, v), h(v)));
        } else {
          vis.edge_not_relaxed(e, g);
        }
        put(color, v, gray_color);
        vis.discover_vertex(v, g);
        Q.push(v);
      } else if (v_color == gray_color) {  // gray_target
        const bool decreased = standin_detail::relax(e, g, weight, predecessor, distance, combine, compare);
        if (decreased) {
          put(cost, v, combine(get(distance, v), h(v)));
          Q.update(v);
          vis.edge_relaxed(e, g);
        } else {
          vis.edge_not_relaxed(e, g);
        }
      } else {  // black_target
        const bool decreased = standin_detail::relax(e, g, weight, predecessor, distance, combine, compare);
        if (decreased) {
          vis.edge_relaxed(e, g);
          put(cost, v, combine(get(distance, v), h(v)));
          Q.push(v);
          put(color, v, gray_color);
          vis.black_target(e, g);
        } else {
          vis.edge_not_relaxed(e, g);
        }
      }
    }
    put(color, u, black_color);
    vis.finish_vertex(u, g);
  }
}
}  // namespace boost
