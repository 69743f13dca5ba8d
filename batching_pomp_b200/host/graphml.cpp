// graphml.cpp — minimal GraphML reader/writer (see graphml.hpp).  A hand-written tokenizer: the
// format needed here is a flat list of <key>, <node>, <data> and <edge> elements.
#include "graphml.hpp"

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

namespace batching_pomp {
namespace util {

namespace {

struct Tag {
  std::string name;  // "node", "/node", "data", ...
  std::unordered_map<std::string, std::string> attr;
  bool selfClosing = false;
};

std::string unescape(const std::string& s) {
  std::string o;
  o.reserve(s.size());
  for (size_t i = 0; i < s.size(); ++i) {
    if (s[i] != '&') { o.push_back(s[i]); continue; }
    size_t e = s.find(';', i);
    if (e == std::string::npos) { o.push_back(s[i]); continue; }
    std::string ent = s.substr(i + 1, e - i - 1);
    if (ent == "amp") o.push_back('&');
    else if (ent == "lt") o.push_back('<');
    else if (ent == "gt") o.push_back('>');
    else if (ent == "quot") o.push_back('"');
    else if (ent == "apos") o.push_back('\'');
    else o += "&" + ent + ";";
    i = e;
  }
  return o;
}

// Parses the inside of <...> (without the brackets).
Tag parseTag(const std::string& body) {
  Tag t;
  size_t i = 0, n = body.size();
  while (i < n && !isspace((unsigned char)body[i]) && body[i] != '/') ++i;
  if (i == 0 && n > 0 && body[0] == '/') {  // closing tag
    size_t j = 1;
    while (j < n && !isspace((unsigned char)body[j])) ++j;
    t.name = body.substr(0, j);
    return t;
  }
  t.name = body.substr(0, i);
  while (i < n) {
    while (i < n && isspace((unsigned char)body[i])) ++i;
    if (i >= n) break;
    if (body[i] == '/') { t.selfClosing = true; ++i; continue; }
    size_t k = i;
    while (k < n && body[k] != '=' && !isspace((unsigned char)body[k])) ++k;
    std::string key = body.substr(i, k - i);
    while (k < n && body[k] != '=') ++k;
    if (k >= n) break;
    ++k;
    while (k < n && isspace((unsigned char)body[k])) ++k;
    if (k >= n || (body[k] != '"' && body[k] != '\'')) throw std::runtime_error("graphml: malformed attribute");
    char q = body[k];
    size_t e = body.find(q, k + 1);
    if (e == std::string::npos) throw std::runtime_error("graphml: unterminated attribute value");
    t.attr[key] = unescape(body.substr(k + 1, e - k - 1));
    i = e + 1;
  }
  return t;
}

}  // namespace

void readGraphml(const std::string& filename, unsigned int dim, std::vector<double>& states,
                 std::vector<std::pair<uint32_t, uint32_t>>& edges) {
  std::ifstream fp(filename.c_str(), std::ios::binary);
  if (!fp) throw std::runtime_error("graphml: cannot open " + filename);
  std::stringstream buf;
  buf << fp.rdbuf();
  const std::string s = buf.str();
  states.clear();
  edges.clear();
  std::unordered_map<std::string, std::string> keyName;  // key id -> attr.name
  std::unordered_map<std::string, uint32_t> nodeIndex;
  std::vector<std::pair<std::string, std::string>> rawEdges;
  bool inNode = false, nodeHasState = false;
  size_t i = 0, n = s.size();
  while (i < n) {
    size_t lt = s.find('<', i);
    if (lt == std::string::npos) break;
    if (s.compare(lt, 4, "<!--") == 0) {
      size_t e = s.find("-->", lt);
      if (e == std::string::npos) throw std::runtime_error("graphml: unterminated comment");
      i = e + 3;
      continue;
    }
    size_t gt = s.find('>', lt);
    if (gt == std::string::npos) throw std::runtime_error("graphml: unterminated tag");
    std::string body = s.substr(lt + 1, gt - lt - 1);
    i = gt + 1;
    if (body.empty() || body[0] == '?' || body[0] == '!') continue;
    Tag t = parseTag(body);
    if (t.name == "key") {
      keyName[t.attr["id"]] = t.attr["attr.name"];
    } else if (t.name == "node") {
      if (nodeIndex.count(t.attr["id"])) throw std::runtime_error("graphml: duplicate node id " + t.attr["id"]);
      nodeIndex[t.attr["id"]] = (uint32_t)(states.size() / dim);
      states.resize(states.size() + dim, 0.0);  // a node without a state keeps a zero state
      inNode = !t.selfClosing;
      nodeHasState = false;
    } else if (t.name == "/node") {
      inNode = false;
    } else if (t.name == "data") {
      std::string text;
      if (!t.selfClosing) {
        size_t e = s.find("</data>", i);
        if (e == std::string::npos) throw std::runtime_error("graphml: unterminated <data>");
        text = unescape(s.substr(i, e - i));
        i = e + 7;
      }
      auto it = keyName.find(t.attr["key"]);
      std::string name = it == keyName.end() ? std::string() : it->second;
      if (name != "state")
        throw std::runtime_error("graphml: property not found: " + (name.empty() ? t.attr["key"] : name));
      if (inNode && !nodeHasState) {
        // RoadmapFromFilePutStateMap put(): `ss >> values[ui]` for ui < dim (RoadmapFromFile.hpp:88-94)
        std::stringstream ss(text);
        double* v = &states[states.size() - dim];
        for (unsigned int ui = 0; ui < dim; ++ui) ss >> v[ui];
        nodeHasState = true;
      }
    } else if (t.name == "edge") {
      rawEdges.emplace_back(t.attr["source"], t.attr["target"]);
    }
  }
  for (auto& e : rawEdges) {
    auto a = nodeIndex.find(e.first), b = nodeIndex.find(e.second);
    if (a == nodeIndex.end() || b == nodeIndex.end()) throw std::runtime_error("graphml: edge names an unknown node");
    edges.emplace_back(a->second, b->second);
  }
}

void writeGraphml(const std::string& filename, unsigned int dim, const std::vector<double>& states,
                  const std::vector<std::pair<uint32_t, uint32_t>>& edges) {
  FILE* f = std::fopen(filename.c_str(), "w");
  if (!f) throw std::runtime_error("graphml: cannot write " + filename);
  std::fprintf(f, "<?xml version='1.0' encoding='utf-8'?>\n"
                  "<graphml xmlns=\"http://graphml.graphdrawing.org/xmlns\" "
                  "xmlns:xsi=\"http://www.w3.org/2001/XMLSchema-instance\" "
                  "xsi:schemaLocation=\"http://graphml.graphdrawing.org/xmlns "
                  "http://graphml.graphdrawing.org/xmlns/1.0/graphml.xsd\">\n"
                  "  <key id=\"d0\" for=\"node\" attr.name=\"state\" attr.type=\"string\" />\n"
                  "  <graph edgedefault=\"undirected\">\n");
  size_t nn = states.size() / dim;
  for (size_t v = 0; v < nn; ++v) {
    std::fprintf(f, "    <node id=\"%zu\">\n      <data key=\"d0\">", v);
    for (unsigned int j = 0; j < dim; ++j) std::fprintf(f, j ? " %.17g" : "%.17g", states[v * dim + j]);
    std::fprintf(f, "</data>\n    </node>\n");
  }
  for (auto& e : edges) std::fprintf(f, "    <edge source=\"%u\" target=\"%u\" />\n", e.first, e.second);
  std::fprintf(f, "  </graph>\n</graphml>\n");
  std::fclose(f);
}

}  // namespace util
}  // namespace batching_pomp
