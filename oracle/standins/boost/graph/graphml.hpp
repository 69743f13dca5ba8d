// Boost.Graph API stand-in (oracle/standins/README.md): read_graphml.  Vertices are created in the order of
// the <node> elements, edges in the order of the <edge> elements (add_edge(source, target)); a <data> element
// of a node whose key names a registered property is handed to that property map as a string.
#pragma once
#include <istream>
#include <iterator>
#include <map>
#include <stdexcept>
#include <string>
#include "boost/graph/adjacency_list.hpp"
#include "boost/property_map/dynamic_property_map.hpp"
namespace boost {
namespace standin_detail {
inline std::string xml_attr(const std::string& tag, const std::string& name) {
  std::size_t p = 0;
  const std::string pat = name + "=";
  while ((p = tag.find(pat, p)) != std::string::npos) {
    if (p > 0 && (tag[p - 1] == ' ' || tag[p - 1] == '\t' || tag[p - 1] == '\n')) {
      const char q = tag[p + pat.size()];
      const std::size_t b = p + pat.size() + 1, e = tag.find(q, b);
      return tag.substr(b, e - b);
    }
    p += pat.size();
  }
  return std::string();
}
}  // namespace standin_detail
template <class Graph> inline void read_graphml(std::istream& in, Graph& g, dynamic_properties& dp, std::size_t = 0) {
  const std::string s((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  std::map<std::string, std::string> key_name;      // key id -> attr.name (node keys)
  std::map<std::string, std::size_t> node_of;       // node id -> vertex
  std::size_t pos = 0;
  long cur_node = -1;
  while ((pos = s.find('<', pos)) != std::string::npos) {
    const std::size_t end = s.find('>', pos);
    if (end == std::string::npos) break;
    const std::string tag = s.substr(pos + 1, end - pos - 1);
    if (tag.compare(0, 4, "key ") == 0) {
      if (standin_detail::xml_attr(tag, "for") != "edge") key_name[standin_detail::xml_attr(tag, "id")] = standin_detail::xml_attr(tag, "attr.name");
    } else if (tag.compare(0, 5, "node ") == 0) {
      const std::size_t v = add_vertex(g);
      node_of[standin_detail::xml_attr(tag, "id")] = v;
      cur_node = tag.back() == '/' ? -1 : (long)v;
    } else if (tag == "/node") {
      cur_node = -1;
    } else if (tag.compare(0, 5, "data ") == 0 && cur_node >= 0) {
      const std::size_t close = s.find("</data>", end);
      const std::string value = s.substr(end + 1, close - end - 1);
      auto it = key_name.find(standin_detail::xml_attr(tag, "key"));
      if (it != key_name.end() && dp.has(it->second)) dp.set(it->second, (std::size_t)cur_node, value);
    } else if (tag.compare(0, 5, "edge ") == 0) {
      auto a = node_of.find(standin_detail::xml_attr(tag, "source")), b = node_of.find(standin_detail::xml_attr(tag, "target"));
      if (a == node_of.end() || b == node_of.end()) throw std::runtime_error("read_graphml: edge references an unknown node");
      add_edge(a->second, b->second, g);
    }
    pos = end + 1;
  }
}
}  // namespace boost
