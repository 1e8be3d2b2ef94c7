// knn_stream_sparse.cu -- second build of knn_stream_kernel: 16 drain warps + 16 expansion-side warps, for target rows with few
// scalar products per panel (recommendation-shaped calls: a user's 200 items against a k-sparse similarity matrix), where
// the sweep of the panel and the selections, not the expansion, are the critical path.  Measured on configs[4]-shaped
// operands (profiles/r02/probe_stream_x.txt): 34.2 ms against 40.0 ms with 8 drain warps and 44.1 ms on the flat engine;
// on configs[1] / configs[3] the 8-warp build is faster (28.2 vs 34.5, 28.9 vs 33.1 ms).
#undef SPY_KS_D_WARPS
#define SPY_KS_D_WARPS 16
#define SPY_KS_TAG b
#include "knn_stream_kernel.cuh"
#include <algorithm>
#include "knn_stream_impl.inc"
