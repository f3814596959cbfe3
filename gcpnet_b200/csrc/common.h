// common.h -- error plumbing shared by the translation units of libgcpnet_b200.so
#pragma once
#include <cuda_runtime.h>

#include <string>

int gcp_fail(const std::string& msg);  // records the message for gcpnet_last_error(), returns 1
void gcp_note_launches(int n);         // bumps the counter behind gcpnet_launch_count()
static inline int fail(const std::string& msg) { return gcp_fail(msg); }
#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t err__ = (expr);                                                                \
    if (err__ != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(err__)); \
  } while (0)

// Optional per-kernel CUDA-event timing (bench.py's roofline leg): when enabled through
// gcpnet_profile_enable(1) every launch of a timed kernel is bracketed by two events on the launching
// stream; gcpnet_profile_read() sums and clears them.  Off by default (and must stay off under capture).
enum GcpTimed { T_EDGE_FWD = 0, T_NODE_FWD = 1, T_NODE_BWD = 2, T_EDGE_BWD = 3, T_COT_REDUCE = 4, T_PARTIAL_REDUCE = 5,
                T_GRAPH_BUILD = 6, T_COUNT = 7 };
bool gcp_profile_on();
void gcp_profile_mark(int which, bool begin, cudaStream_t st);
struct GcpTimedScope {
  int which; cudaStream_t st; bool on;
  GcpTimedScope(int w, cudaStream_t s) : which(w), st(s), on(gcp_profile_on()) { if (on) gcp_profile_mark(which, true, st); }
  ~GcpTimedScope() { if (on) gcp_profile_mark(which, false, st); }
};
