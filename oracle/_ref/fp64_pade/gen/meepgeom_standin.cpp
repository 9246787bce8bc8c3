#include "meepgeom.hpp"
namespace meep_geom {
double fragment_stats::resolution = 0;
bool fragment_stats::split_chunks_evenly = true;
} // namespace meep_geom
