// graphml.cpp — minimal GraphML reader/writer (see graphml.hpp).  A hand-written tokenizer: the
// format needed here is a flat list of <key>, <node>, <data> and <edge> elements.
#include "graphml.hpp"

#include <cctype>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <unordered_map>

#include <sys/stat.h>

namespace batching_pomp {
namespace util {

namespace {

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

}  // namespace

namespace {

// Value of attribute `name` inside a tag body [b, e) as a (pointer, length) view; false if absent.
bool attrView(const char* b, const char* e, const char* name, const char*& vb, size_t& vn) {
  const size_t ln = std::strlen(name);
  for (const char* p = b; p + ln + 2 < e; ++p) {
    if ((p == b || isspace((unsigned char)p[-1])) && std::strncmp(p, name, ln) == 0) {
      const char* q = p + ln;
      while (q < e && isspace((unsigned char)*q)) ++q;
      if (q >= e || *q != '=') continue;
      ++q;
      while (q < e && isspace((unsigned char)*q)) ++q;
      if (q >= e || (*q != '"' && *q != '\'')) throw std::runtime_error("graphml: malformed attribute");
      const char quote = *q++;
      const char* r = static_cast<const char*>(std::memchr(q, quote, (size_t)(e - q)));
      if (!r) throw std::runtime_error("graphml: unterminated attribute value");
      vb = q;
      vn = (size_t)(r - q);
      return true;
    }
  }
  return false;
}

bool tagIs(const char* b, const char* e, const char* name) {
  const size_t ln = std::strlen(name);
  return (size_t)(e - b) >= ln && std::strncmp(b, name, ln) == 0 &&
         ((size_t)(e - b) == ln || isspace((unsigned char)b[ln]) || b[ln] == '/' );
}

}  // namespace

// Single pass over the file buffer; numbers are parsed in place with std::from_chars (same correctly rounded
// decimal-to-double conversion as the `ss >> values[ui]` of RoadmapFromFile.hpp:91-94).  A 10^6-node
// 7-D roadmap (200 MB) loads in well under a second.
void readGraphml(const std::string& filename, unsigned int dim, std::vector<double>& states,
                 std::vector<std::pair<uint32_t, uint32_t>>& edges) {
  FILE* fp = std::fopen(filename.c_str(), "rb");
  if (!fp) throw std::runtime_error("graphml: cannot open " + filename);
  std::string buf;
  std::fseek(fp, 0, SEEK_END);
  long sz = std::ftell(fp);
  std::fseek(fp, 0, SEEK_SET);
  buf.resize(sz > 0 ? (size_t)sz : 0);
  size_t got = buf.empty() ? 0 : std::fread(&buf[0], 1, buf.size(), fp);
  std::fclose(fp);
  buf.resize(got);
  buf.push_back('\0');  // strtod sentinel
  states.clear();
  edges.clear();
  std::unordered_map<std::string, std::string> keyName;  // key id -> attr.name
  std::string stateKey;                                   // id of the key whose attr.name is "state"
  std::vector<std::pair<const char*, size_t>> nodeIds;   // views into buf, in <node> order
  std::vector<std::pair<std::string, std::string>> rawEdges;
  bool inNode = false, nodeHasState = false;
  const char* p = buf.data();
  const char* const end = buf.data() + got;
  while (p < end) {
    const char* lt = static_cast<const char*>(std::memchr(p, '<', (size_t)(end - p)));
    if (!lt) break;
    if (end - lt >= 4 && std::strncmp(lt, "<!--", 4) == 0) {
      const char* e = std::strstr(lt, "-->");
      if (!e) throw std::runtime_error("graphml: unterminated comment");
      p = e + 3;
      continue;
    }
    const char* gt = static_cast<const char*>(std::memchr(lt, '>', (size_t)(end - lt)));
    if (!gt) throw std::runtime_error("graphml: unterminated tag");
    const char* b = lt + 1;
    const char* e = gt;
    p = gt + 1;
    if (b >= e || *b == '?' || *b == '!') continue;
    const bool selfClosing = e[-1] == '/';
    if (tagIs(b, e, "data")) {
      const char* tb = p;
      const char* te = p;
      if (!selfClosing) {
        te = std::strstr(p, "</data>");
        if (!te) throw std::runtime_error("graphml: unterminated <data>");
        p = te + 7;
      }
      const char* kb = nullptr;
      size_t kn = 0;
      bool isState = attrView(b, e, "key", kb, kn) && kn == stateKey.size() && !stateKey.empty() &&
                     std::strncmp(kb, stateKey.data(), kn) == 0;
      if (!isState) {
        std::string kid = kb ? std::string(kb, kn) : std::string();
        auto it = keyName.find(kid);
        std::string name = it == keyName.end() ? std::string() : it->second;
        if (name != "state") throw std::runtime_error("graphml: property not found: " + (name.empty() ? kid : name));
      }
      if (inNode && !nodeHasState) {
        // RoadmapFromFilePutStateMap put(): `ss >> values[ui]` for ui < dim (RoadmapFromFile.hpp:88-94)
        double* v = &states[states.size() - dim];
        const char* q = tb;
        for (unsigned int ui = 0; ui < dim; ++ui) {
          while (q < te && isspace((unsigned char)*q)) ++q;
          if (q >= te) break;  // fewer than dim numbers: the rest stays as allocated (zero here)
          if (*q == '+') ++q;  // from_chars does not take a leading '+'
          double val = 0.0;
          auto res = std::from_chars(q, te, val);  // correctly rounded, like strtod / operator>>
          if (res.ec != std::errc()) {             // inf / nan spellings etc.: fall back to strtod
            char* endp = nullptr;
            val = std::strtod(q, &endp);
            if (endp == q || endp > te) break;
            res.ptr = endp;
          }
          v[ui] = val;
          q = res.ptr;
        }
        nodeHasState = true;
      }
    } else if (tagIs(b, e, "node")) {
      const char* ib = nullptr;
      size_t in = 0;
      if (!attrView(b, e, "id", ib, in)) throw std::runtime_error("graphml: <node> without id");
      nodeIds.emplace_back(ib, in);
      states.resize(states.size() + dim, 0.0);  // a node without a state keeps a zero state
      inNode = !selfClosing;
      nodeHasState = false;
    } else if (tagIs(b, e, "/node")) {
      inNode = false;
    } else if (tagIs(b, e, "key")) {
      const char *ib = nullptr, *nb = nullptr;
      size_t in = 0, nn = 0;
      if (attrView(b, e, "id", ib, in)) {
        std::string name = attrView(b, e, "attr.name", nb, nn) ? unescape(std::string(nb, nn)) : std::string();
        keyName[std::string(ib, in)] = name;
        if (name == "state") stateKey.assign(ib, in);
      }
    } else if (tagIs(b, e, "edge")) {
      const char *sb = nullptr, *tb2 = nullptr;
      size_t sn = 0, tn = 0;
      if (!attrView(b, e, "source", sb, sn) || !attrView(b, e, "target", tb2, tn))
        throw std::runtime_error("graphml: <edge> without source/target");
      rawEdges.emplace_back(std::string(sb, sn), std::string(tb2, tn));
    }
  }
  if (!rawEdges.empty()) {
    std::unordered_map<std::string, uint32_t> nodeIndex;
    nodeIndex.reserve(nodeIds.size() * 2);
    for (size_t i = 0; i < nodeIds.size(); ++i) {
      auto ins = nodeIndex.emplace(std::string(nodeIds[i].first, nodeIds[i].second), (uint32_t)i);
      if (!ins.second) throw std::runtime_error("graphml: duplicate node id " + ins.first->first);
    }
    for (auto& ed : rawEdges) {
      auto a = nodeIndex.find(ed.first), b2 = nodeIndex.find(ed.second);
      if (a == nodeIndex.end() || b2 == nodeIndex.end())
        throw std::runtime_error("graphml: edge names an unknown node");
      edges.emplace_back(a->second, b2->second);
    }
  }
}

namespace {
struct CacheHeader {
  char magic[8];          // "BPOMPRM1"
  uint32_t dim;
  uint32_t reserved;
  uint64_t n, nedges;
  uint64_t src_size;      // the GraphML file the cache was made from
  int64_t src_mtime_ns;
  uint64_t checksum;      // FNV-1a over the payload
};
uint64_t fnv1a(const void* data, size_t n, uint64_t h) {
  const unsigned char* b = static_cast<const unsigned char*>(data);
  for (size_t i = 0; i < n; ++i) {
    h ^= b[i];
    h *= 1099511628211ull;
  }
  return h;
}
bool sourceStamp(const std::string& filename, uint64_t& size, int64_t& mtime_ns) {
  struct stat st;
  if (stat(filename.c_str(), &st) != 0) return false;
  size = (uint64_t)st.st_size;
  mtime_ns = (int64_t)st.st_mtim.tv_sec * 1000000000ll + (int64_t)st.st_mtim.tv_nsec;
  return true;
}
}  // namespace

bool readGraphmlCached(const std::string& filename, unsigned int dim, std::vector<double>& states,
                       std::vector<std::pair<uint32_t, uint32_t>>& edges) {
  const char* env = std::getenv("BPOMP_ROADMAP_CACHE");
  const bool enabled = !(env && env[0] == '0');
  const std::string cache = filename + ".bpomp-cache";
  uint64_t size = 0;
  int64_t mtime = 0;
  const bool stamped = enabled && sourceStamp(filename, size, mtime);
  if (stamped) {
    if (FILE* f = std::fopen(cache.c_str(), "rb")) {
      CacheHeader h;
      bool ok = std::fread(&h, sizeof(h), 1, f) == 1 && std::memcmp(h.magic, "BPOMPRM1", 8) == 0 && h.dim == dim &&
                h.src_size == size && h.src_mtime_ns == mtime && h.n < (1ull << 32) && h.nedges < (1ull << 32);
      if (ok) {
        std::vector<double> s((size_t)h.n * dim);
        std::vector<std::pair<uint32_t, uint32_t>> e((size_t)h.nedges);
        ok = (s.empty() || std::fread(s.data(), sizeof(double), s.size(), f) == s.size()) &&
             (e.empty() || std::fread(e.data(), sizeof(e[0]), e.size(), f) == e.size());
        if (ok) {
          uint64_t c = fnv1a(s.data(), s.size() * sizeof(double), 1469598103934665603ull);
          c = fnv1a(e.data(), e.size() * sizeof(e[0]), c);
          ok = c == h.checksum;
        }
        if (ok) {
          std::fclose(f);
          states.swap(s);
          edges.swap(e);
          return true;
        }
      }
      std::fclose(f);
    }
  }
  readGraphml(filename, dim, states, edges);
  if (stamped) {  // best effort: a read-only directory just means no cache
    const std::string tmp = cache + ".tmp";
    if (FILE* f = std::fopen(tmp.c_str(), "wb")) {
      CacheHeader h;
      std::memset(&h, 0, sizeof(h));
      std::memcpy(h.magic, "BPOMPRM1", 8);
      h.dim = dim;
      h.n = states.size() / dim;
      h.nedges = edges.size();
      h.src_size = size;
      h.src_mtime_ns = mtime;
      h.checksum = fnv1a(states.data(), states.size() * sizeof(double), 1469598103934665603ull);
      h.checksum = fnv1a(edges.data(), edges.size() * sizeof(edges[0]), h.checksum);
      bool ok = std::fwrite(&h, sizeof(h), 1, f) == 1 &&
                (states.empty() || std::fwrite(states.data(), sizeof(double), states.size(), f) == states.size()) &&
                (edges.empty() || std::fwrite(edges.data(), sizeof(edges[0]), edges.size(), f) == edges.size());
      ok = (std::fclose(f) == 0) && ok;
      if (!ok || std::rename(tmp.c_str(), cache.c_str()) != 0) std::remove(tmp.c_str());
    }
  }
  return false;
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
