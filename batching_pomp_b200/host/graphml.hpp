// graphml.hpp — roadmap input in the reference's boundary format.
// Replaces boost::read_graphml as used by util::RoadmapFromFile
// (/root/reference/include/batching_pomp/util/RoadmapFromFile.hpp:82-95, 136-160): the only
// registered property is the vertex attribute "state", a whitespace-separated list of `dim` doubles;
// vertices are numbered in <node> order; <edge> elements are kept for single batching.
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace batching_pomp {
namespace util {

/// Throws std::runtime_error on a missing file, malformed XML, a <data> element that refers to a
/// property other than "state" (boost::read_graphml raises property_not_found there) or an edge
/// that names an unknown node.
void readGraphml(const std::string& filename, unsigned int dim, std::vector<double>& states,
                 std::vector<std::pair<uint32_t, uint32_t>>& edges);

/// Writes a roadmap the way networkx writes it (one string attribute "state" per node); doubles are
/// printed with 17 significant digits so that `ss >> double` restores them exactly.
void writeGraphml(const std::string& filename, unsigned int dim, const std::vector<double>& states,
                  const std::vector<std::pair<uint32_t, uint32_t>>& edges);

}  // namespace util
}  // namespace batching_pomp
