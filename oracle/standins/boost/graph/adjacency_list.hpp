// Boost.Graph API stand-in (oracle/standins/README.md): adjacency_list<vecS, vecS, undirectedS, VP, EP> only.
// Behaviour restated from the library's documentation: vertex descriptor = index; each vertex
// keeps its out-edges in insertion order and add_edge(u,v) appends to both u's and v's lists (parallel edges
// allowed); the global edge list is in creation order; edge(u,v,g) scans u's out-edges; clear_vertex(v)
// removes every edge incident to v from both endpoints, keeping the relative order of the others; an edge
// descriptor carries (source, target, property pointer), compares equal on the property pointer and is
// ordered by it; the descriptor returned by add_edge(u,v) has source u, target v, the ones obtained from
// out_edges(x) have source x.
#pragma once
// (the real header pulls in most of the standard library; the reference relies on that for <set>, <cstring>, ...)
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>
#include "boost/graph/properties.hpp"
namespace boost {
struct vecS {};
struct undirectedS {};
struct undirected_tag {};
struct no_property {};
namespace standin_detail {
// Boost orders edge descriptors by the ADDRESS of their property object (edge_desc_impl::operator<), i.e. by
// whatever the allocator returned.  To keep runs reproducible the stand-in orders them by a creation serial
// instead — the address order of an allocator that hands out increasing addresses and never reuses one.
template <class EP> struct edge_serial_prop : EP {
  unsigned long long serial = 0;
};
template <class EP> struct edge_desc {
  std::size_t m_source = 0, m_target = 0;
  EP* m_prop = nullptr;
  bool operator==(const edge_desc& o) const { return m_prop == o.m_prop; }
  bool operator!=(const edge_desc& o) const { return m_prop != o.m_prop; }
  bool operator<(const edge_desc& o) const {
    return static_cast<const edge_serial_prop<EP>*>(m_prop)->serial < static_cast<const edge_serial_prop<EP>*>(o.m_prop)->serial;
  }
};
struct counting_iterator {
  typedef std::forward_iterator_tag iterator_category;
  typedef std::size_t value_type;
  typedef std::ptrdiff_t difference_type;
  typedef const std::size_t* pointer;
  typedef std::size_t reference;
  std::size_t i = 0;
  counting_iterator() = default;
  explicit counting_iterator(std::size_t v) : i(v) {}
  std::size_t operator*() const { return i; }
  counting_iterator& operator++() { ++i; return *this; }
  counting_iterator operator++(int) { counting_iterator t = *this; ++i; return t; }
  bool operator==(const counting_iterator& o) const { return i == o.i; }
  bool operator!=(const counting_iterator& o) const { return i != o.i; }
};
}  // namespace standin_detail

template <class OutEdgeListS, class VertexListS, class DirectedS, class VP = no_property, class EP = no_property>
class adjacency_list {
public:
  typedef std::size_t vertex_descriptor;
  typedef standin_detail::edge_desc<EP> edge_descriptor;
  typedef standin_detail::counting_iterator vertex_iterator;
  typedef undirected_tag directed_category;
  typedef VP vertex_bundled;
  typedef EP edge_bundled;
  struct stored_edge {  // an entry of a vertex's out-edge list
    std::size_t target;
    typename std::list<std::pair<std::pair<std::size_t, std::size_t>, standin_detail::edge_serial_prop<EP>>>::iterator it;
  };
  typedef std::list<std::pair<std::pair<std::size_t, std::size_t>, standin_detail::edge_serial_prop<EP>>> edge_list_t;
  unsigned long long next_serial_ = 0;
  class out_edge_iterator {
  public:
    typedef std::forward_iterator_tag iterator_category;
    typedef edge_descriptor value_type;
    typedef std::ptrdiff_t difference_type;
    typedef const edge_descriptor* pointer;
    typedef edge_descriptor reference;
    out_edge_iterator() = default;
    out_edge_iterator(std::size_t u, const std::vector<stored_edge>* v, std::size_t i) : u_(u), v_(v), i_(i) {}
    edge_descriptor operator*() const {
      edge_descriptor e;
      e.m_source = u_;
      e.m_target = (*v_)[i_].target;
      e.m_prop = &(*v_)[i_].it->second;
      return e;
    }
    out_edge_iterator& operator++() { ++i_; return *this; }
    bool operator==(const out_edge_iterator& o) const { return i_ == o.i_; }
    bool operator!=(const out_edge_iterator& o) const { return i_ != o.i_; }
  private:
    std::size_t u_ = 0;
    const std::vector<stored_edge>* v_ = nullptr;
    std::size_t i_ = 0;
  };
  class edge_iterator {
  public:
    typedef std::forward_iterator_tag iterator_category;
    typedef edge_descriptor value_type;
    typedef std::ptrdiff_t difference_type;
    typedef const edge_descriptor* pointer;
    typedef edge_descriptor reference;
    edge_iterator() = default;
    explicit edge_iterator(typename edge_list_t::iterator it) : it_(it) {}
    edge_descriptor operator*() const {
      edge_descriptor e;
      e.m_source = it_->first.first;
      e.m_target = it_->first.second;
      e.m_prop = &it_->second;
      return e;
    }
    edge_iterator& operator++() { ++it_; return *this; }
    bool operator==(const edge_iterator& o) const { return it_ == o.it_; }
    bool operator!=(const edge_iterator& o) const { return it_ != o.it_; }
  private:
    typename edge_list_t::iterator it_;
  };

  adjacency_list() = default;
  adjacency_list(const adjacency_list& o) { *this = o; }
  // copy: the vertices, then the edges in creation order (as Boost's copy does through add_edge)
  adjacency_list& operator=(const adjacency_list& o) {
    if (this == &o) return *this;
    vprops_ = o.vprops_;
    out_.assign(vprops_.size(), std::vector<stored_edge>());
    edges_.clear();
    for (auto it = o.edges_.begin(); it != o.edges_.end(); ++it) {
      edges_.push_back(*it);
      auto me = std::prev(edges_.end());
      me->second.serial = next_serial_++;
      out_[it->first.first].push_back({it->first.second, me});
      out_[it->first.second].push_back({it->first.first, me});
    }
    return *this;
  }

  VP& operator[](vertex_descriptor v) { return vprops_[v]; }
  const VP& operator[](vertex_descriptor v) const { return vprops_[v]; }
  EP& operator[](const edge_descriptor& e) { return *e.m_prop; }
  const EP& operator[](const edge_descriptor& e) const { return *e.m_prop; }

  std::vector<VP> vprops_;
  std::vector<std::vector<stored_edge>> out_;
  mutable edge_list_t edges_;
};

template <class G> struct graph_traits {
  typedef typename G::vertex_descriptor vertex_descriptor;
  typedef typename G::edge_descriptor edge_descriptor;
  typedef typename G::vertex_iterator vertex_iterator;
  typedef typename G::edge_iterator edge_iterator;
  typedef typename G::out_edge_iterator out_edge_iterator;
  typedef typename G::directed_category directed_category;
};

#define BP_ADJ_TMPL template <class O, class V, class D, class VP, class EP>
#define BP_ADJ adjacency_list<O, V, D, VP, EP>
BP_ADJ_TMPL inline std::size_t add_vertex(BP_ADJ& g) {
  g.vprops_.emplace_back();
  g.out_.emplace_back();
  return g.vprops_.size() - 1;
}
BP_ADJ_TMPL inline std::size_t num_vertices(const BP_ADJ& g) { return g.vprops_.size(); }
BP_ADJ_TMPL inline std::size_t num_edges(const BP_ADJ& g) { return g.edges_.size(); }
BP_ADJ_TMPL inline std::pair<typename BP_ADJ::vertex_iterator, typename BP_ADJ::vertex_iterator> vertices(const BP_ADJ& g) {
  return std::make_pair(typename BP_ADJ::vertex_iterator(0), typename BP_ADJ::vertex_iterator(g.vprops_.size()));
}
BP_ADJ_TMPL inline std::pair<typename BP_ADJ::edge_iterator, typename BP_ADJ::edge_iterator> edges(const BP_ADJ& g) {
  return std::make_pair(typename BP_ADJ::edge_iterator(g.edges_.begin()), typename BP_ADJ::edge_iterator(g.edges_.end()));
}
BP_ADJ_TMPL inline std::pair<typename BP_ADJ::out_edge_iterator, typename BP_ADJ::out_edge_iterator> out_edges(std::size_t u, const BP_ADJ& g) {
  return std::make_pair(typename BP_ADJ::out_edge_iterator(u, &g.out_[u], 0),
                        typename BP_ADJ::out_edge_iterator(u, &g.out_[u], g.out_[u].size()));
}
BP_ADJ_TMPL inline std::pair<typename BP_ADJ::edge_descriptor, bool> add_edge(std::size_t u, std::size_t v, BP_ADJ& g) {
  g.edges_.push_back(std::make_pair(std::make_pair(u, v), standin_detail::edge_serial_prop<EP>()));
  auto it = std::prev(g.edges_.end());
  it->second.serial = g.next_serial_++;
  g.out_[u].push_back({v, it});
  g.out_[v].push_back({u, it});  // (a self-loop would be stored twice, as in Boost)
  typename BP_ADJ::edge_descriptor e;
  e.m_source = u;
  e.m_target = v;
  e.m_prop = &it->second;
  return std::make_pair(e, true);
}
BP_ADJ_TMPL inline std::pair<typename BP_ADJ::edge_descriptor, bool> edge(std::size_t u, std::size_t v, const BP_ADJ& g) {
  typename BP_ADJ::edge_descriptor e;
  for (const auto& se : g.out_[u])
    if (se.target == v) {
      e.m_source = u;
      e.m_target = v;
      e.m_prop = &se.it->second;
      return std::make_pair(e, true);
    }
  return std::make_pair(e, false);
}
BP_ADJ_TMPL inline void clear_vertex(std::size_t v, BP_ADJ& g) {
  for (const auto& se : g.out_[v]) {
    if (se.target != v) {
      auto& other = g.out_[se.target];
      for (std::size_t i = 0; i < other.size(); ++i)
        if (other[i].it == se.it) {
          other.erase(other.begin() + (std::ptrdiff_t)i);
          break;
        }
    }
  }
  std::vector<typename BP_ADJ::edge_list_t::iterator> dead;
  for (const auto& se : g.out_[v]) {
    bool seen = false;
    for (auto& d : dead) seen = seen || d == se.it;
    if (!seen) dead.push_back(se.it);
  }
  g.out_[v].clear();
  for (auto& d : dead) g.edges_.erase(d);
}
template <class EP, class G> inline std::size_t source(const standin_detail::edge_desc<EP>& e, const G&) { return e.m_source; }
template <class EP, class G> inline std::size_t target(const standin_detail::edge_desc<EP>& e, const G&) { return e.m_target; }

// bundled property maps: get(&VP::member, g) / get(&EP::member, g), and vertex_index
template <class G, class Bundle, class T, bool IsVertex> class bundle_property_map {
public:
  typedef typename std::conditional<IsVertex, typename G::vertex_descriptor, typename G::edge_descriptor>::type key_type;
  typedef T value_type;
  typedef T& reference;
  typedef lvalue_property_map_tag category;
  bundle_property_map() : g_(nullptr), pm_(nullptr) {}
  bundle_property_map(G* g, T Bundle::*pm) : g_(g), pm_(pm) {}
  reference operator[](const key_type& k) const { return (*g_)[k].*pm_; }
private:
  G* g_;
  T Bundle::*pm_;
};
template <class G, class B, class T, bool V> inline T& get(const bundle_property_map<G, B, T, V>& pm, const typename bundle_property_map<G, B, T, V>::key_type& k) { return pm[k]; }
template <class G, class B, class T, bool V, class U> inline void put(const bundle_property_map<G, B, T, V>& pm, const typename bundle_property_map<G, B, T, V>::key_type& k, const U& v) { pm[k] = v; }
struct identity_index_map {
  typedef std::size_t key_type;
  typedef std::size_t value_type;
  typedef std::size_t reference;
  typedef readable_property_map_tag category;
  std::size_t operator[](std::size_t v) const { return v; }
};
inline std::size_t get(const identity_index_map&, std::size_t v) { return v; }

template <class G, class Tag> struct property_map;
template <class O, class V, class D, class VP, class EP, class T> struct property_map<BP_ADJ, T VP::*> {
  typedef bundle_property_map<BP_ADJ, VP, T, true> type;
  typedef type const_type;
};
template <class O, class V, class D, class VP, class EP> struct property_map<BP_ADJ, vertex_index_t> {
  typedef identity_index_map type;
  typedef type const_type;
};
// an edge-bundle member pointer (VP and EP are different types in every use)
template <class G, class T, class EPx> struct edge_member_map { typedef bundle_property_map<G, EPx, T, false> type; };
template <class O, class V, class D, class VP, class EP, class T>
struct property_map<BP_ADJ, T EP::*> {
  typedef bundle_property_map<BP_ADJ, EP, T, false> type;
  typedef type const_type;
};
BP_ADJ_TMPL inline identity_index_map get(vertex_index_t, const BP_ADJ&) { return identity_index_map(); }
template <class O, class V, class D, class VP, class EP, class T>
inline bundle_property_map<BP_ADJ, VP, T, true> get(T VP::*pm, BP_ADJ& g) { return bundle_property_map<BP_ADJ, VP, T, true>(&g, pm); }
template <class O, class V, class D, class VP, class EP, class T>
inline bundle_property_map<BP_ADJ, EP, T, false> get(T EP::*pm, BP_ADJ& g) { return bundle_property_map<BP_ADJ, EP, T, false>(&g, pm); }
#undef BP_ADJ_TMPL
#undef BP_ADJ
}  // namespace boost
