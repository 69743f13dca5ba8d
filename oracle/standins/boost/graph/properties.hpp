// Boost.Graph API stand-in (oracle/standins/README.md)
#pragma once
#include <map>
#include <utility>
namespace boost {
enum default_color_type { white_color, gray_color, green_color, red_color, black_color };
template <class C> struct color_traits {
  static default_color_type white() { return white_color; }
  static default_color_type gray() { return gray_color; }
  static default_color_type black() { return black_color; }
};
enum vertex_index_t { vertex_index };
struct readable_property_map_tag {};
struct writable_property_map_tag {};
struct read_write_property_map_tag : readable_property_map_tag, writable_property_map_tag {};
struct lvalue_property_map_tag : read_write_property_map_tag {};
template <class PMap> struct property_traits {
  typedef typename PMap::key_type key_type;
  typedef typename PMap::value_type value_type;
  typedef typename PMap::reference reference;
  typedef typename PMap::category category;
};
// boost::tie for pairs
template <class A, class B> struct tie_holder {
  A& a; B& b;
  tie_holder(A& a_, B& b_) : a(a_), b(b_) {}
  template <class U, class V> tie_holder& operator=(const std::pair<U, V>& p) { a = p.first; b = p.second; return *this; }
};
template <class A, class B> inline tie_holder<A, B> tie(A& a, B& b) { return tie_holder<A, B>(a, b); }
// associative_property_map over a std::map: get() default-constructs missing entries (operator[])
template <class Map> class associative_property_map {
public:
  typedef typename Map::key_type key_type;
  typedef typename Map::mapped_type value_type;
  typedef value_type& reference;
  typedef lvalue_property_map_tag category;
  associative_property_map() : m_(nullptr) {}
  explicit associative_property_map(Map& m) : m_(&m) {}
  reference operator[](const key_type& k) const { return (*m_)[k]; }
private:
  Map* m_;
};
template <class Map> inline associative_property_map<Map> make_assoc_property_map(Map& m) { return associative_property_map<Map>(m); }
template <class Map> inline typename Map::mapped_type& get(const associative_property_map<Map>& pm, const typename Map::key_type& k) { return pm[k]; }
template <class Map, class V> inline void put(const associative_property_map<Map>& pm, const typename Map::key_type& k, const V& v) { pm[k] = v; }
}  // namespace boost
