// Boost.Graph API stand-in (oracle/standins/README.md): included by the reference, nothing of it is used.
#pragma once
#include "boost/graph/adjacency_list.hpp"
