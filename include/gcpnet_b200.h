/* gcpnet_b200.h -- C ABI of the B200-native GCPNet message-passing layer.
 *
 * The reference (BioinfoMachineLearning/GCPNet @ 172733b) is pure Python: its "operator interface"
 * for this path is the nn.Module GCPInteractions (src/models/components/gcpnet.py:963-1262) and the
 * torch_scatter.scatter calls under it (gcpnet.py:946; src/models/components/__init__.py:316).
 * This header is the boundary a binding for that path links against: plain pointers and sizes,
 * device memory owned by the caller, every call enqueues work on the caller's CUDA stream and
 * returns immediately (no allocation, no synchronisation -> safe inside CUDA-graph capture).
 * gcpnet_b200/interactions.py (ctypes) is the reference-facing binding; INTEGRATION.md shows it.
 *
 * All feature tensors are fp32, row-major, contiguous:
 *   h[N][s]  chi[N][v][3]  e[E][se]  xi[E][ve][3]  frames[E][3][3]  pos[N][3]
 *   edge_index int64 [2][E] exactly as the reference receives it (gcpnet.py:1165), any order,
 *   self loops / duplicate edges / isolated nodes allowed.
 * Return value: 0 on success, non-zero on error (message via gcpnet_last_error()).
 */
#ifndef GCPNET_B200_H
#define GCPNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCPNET_MAX_MESSAGE_LAYERS 12

/* scalar / vector nonlinearities: src/models/__init__.py:41-57 (get_nonlinearity) */
enum gcpnet_act { GCPNET_ACT_NONE = 0, GCPNET_ACT_RELU = 1, GCPNET_ACT_LEAKYRELU = 2, GCPNET_ACT_SILU = 3,
                  GCPNET_ACT_SIGMOID = 4, GCPNET_ACT_SELU = 5 };

/* One GCP2 module (gcpnet.py:252-468, vector_gate path).  Weights keep the reference's nn.Linear
 * layouts and state_dict names: vector_down.weight[hd][vi], vector_down_frames.weight[3][vi],
 * scalar_out.weight[so][si+hd+9] (+bias[so]), vector_up.weight[vo][hd],
 * vector_out_scale.weight[vo][so] (+bias[vo]).  grad_off[i] = offset (floats) of the i-th block's
 * gradient inside the layer's flat gradient, in the order listed. */
/* gcpnet_gcp2.flags (the GCP-Baseline variants the CPD decoder is built with, gcpnet_cpd_module.py:95-97):
 *   NO_FRAMES  ablate_frame_updates=True (gcpnet.py:302-309,424-437): no vector_down_frames, scalar_out reads [s | norms] only;
 *   NO_GATE    vector_gate=False with an identity vector nonlinearity (gcpnet.py:344-350): V' = vector_up(H) (+ V), no
 *              vector_out_scale.  The pointers (and gradient slots) of the absent parameters are ignored.
 * Both run the FFMA tile kernels. */
#define GCPNET_GCP2_NO_FRAMES 1
#define GCPNET_GCP2_NO_GATE 2

typedef struct gcpnet_gcp2 {
  int32_t si, vi, so, vo, hd;
  int32_t act_s, act_v, vector_residual;
  const float* vector_down;
  const float* vector_down_frames;
  const float* scalar_out_w;
  const float* scalar_out_b;
  const float* vector_up;
  const float* vector_out_scale_w;
  const float* vector_out_scale_b;
  int32_t grad_off[7]; /* vector_down, vector_down_frames, scalar_out_w, scalar_out_b, vector_up, vector_out_scale_w, vector_out_scale_b */
  int32_t flags;       /* GCPNET_GCP2_NO_FRAMES | GCPNET_GCP2_NO_GATE */
} gcpnet_gcp2;

/* One GCPInteractions layer (gcpnet.py:963-1063): message stack, two LayerNorms, two feed-forward
 * GCPs, optional position-update GCP.  Flat gradient layout = [message_fusion.* | everything else];
 * n_edge_params / n_node_params are the two block sizes in floats. */
typedef struct gcpnet_layer {
  int32_t s, v, se, ve;            /* node (s, v) and edge (se, ve) hidden dims */
  int32_t num_message_layers;      /* mp_cfg.num_message_layers */
  int32_t residual_messages;       /* mp_cfg.use_residual_message_gcp */
  int32_t reduce_mean;             /* 1: "mean" (default), 0: "add"/"sum" (gcpnet.py:984) */
  int32_t enable_e3;               /* cfg.enable_e3_equivariance (edge path only) */
  int32_t has_pos;                 /* updating_node_positions */
  int32_t training;                /* 1: dropout active (GCPDropout, comp/__init__.py:97-135) */
  float slope;                     /* leaky-relu slope */
  float ln_eps, vn_eps;            /* nn.LayerNorm eps (1e-5), GCPLayerNorm vector eps (1e-8) */
  float pos_weight;                /* cfg.node_positions_weight */
  float p_drop;                    /* dropout probability */
  uint64_t seed;                   /* dropout stream seed */
  const int64_t* rng_counter;      /* device int64, advanced by the caller once per training forward */
  gcpnet_gcp2 message[GCPNET_MAX_MESSAGE_LAYERS]; /* interaction.message_fusion.{k} */
  gcpnet_gcp2 ff0, ff1;            /* feedforward_network.{0,1} */
  gcpnet_gcp2 pos_update;          /* node_position_update_network.0 (if has_pos) */
  const float *ln0_w, *ln0_b, *ln1_w, *ln1_b; /* gcp_norm.{0,1}.scalar_norm.{weight,bias} */
  int32_t ln_grad_off[4];
  int32_t n_edge_params, n_node_params;
  int32_t pre_norm;                /* layer_cfg.pre_norm (gcpnet.py:1188-1189,1223-1224,1245): gcp_norm.0 before the message
                                      passing, gcp_norm.1 after the first residual, none at the end */
  int32_t autoregressive;          /* 1: the graph carries the gather views of gcpnet_graph_build_autoregressive; 2: aggregate_with_row
                                      (gcpnet.py:946): graph built on the flipped edge_index, gsrc = dst / gdst = src swap the
                                      ends back, num_gather_rows = 0.  Both run the FFMA edge kernels. */
  /* GCPMessagePassing(use_scalar_message_attention=True) (gcpnet.py:893-897,931-934; the EQ / AR layers): the scalar
   * messages are scaled by sigmoid(attn_w . m_s + attn_b) before the aggregation.  NULL = off.  attn_grad_off = offsets
   * of the two gradients inside the edge block of the flat gradient (they count in n_edge_params).  FFMA edge kernels. */
  const float* attn_w;             /* interaction.scalar_message_attention.0.weight [1][s] */
  const float* attn_b;             /* interaction.scalar_message_attention.0.bias [1] */
  int32_t attn_grad_off[2];
} gcpnet_layer;

/* Destination- and source-sorted views of one graph batch, built by gcpnet_graph_build and shared
 * by every layer that sees the same (edge_index, frames). */
typedef struct gcpnet_graph {
  int64_t num_nodes, num_edges;
  const int32_t* perm;     /* [E] sorted position -> caller's edge id (stable sort by destination) */
  const int32_t* src;      /* [E] source node of sorted position */
  const int32_t* dst;      /* [E] destination node of sorted position (non-decreasing) */
  const int32_t* dst_ptr;  /* [N+1] CSR pointer over `dst` */
  const int32_t* src_pos;  /* [E] sorted positions grouped by source node (stable) */
  const int32_t* src_ptr;  /* [N+1] CSR pointer over src_pos */
  const float* fbar;       /* [N][9] mean frame over the edges leaving each node (0 if none): node-side scalarize of the
                              feed-forward GCPs (comp/__init__.py:316-323); under a node mask the subgraph version */
  /* ---- optional views (NULL / 0 when absent) ---- */
  const float* fbar_pos;   /* [N][9] node mask: mean frame for the position-update GCP (full graph, masked edges give zero
                              frames but count, comp/__init__.py:294-300,316-323); NULL = fbar */
  const uint8_t* node_mask; /* [N] 1 = node takes part in the update (gcpnet.py:1202-1217,1249-1251); NULL = all */
  const int32_t* gsrc;     /* [E] autoregressive: row of the gather table (2 * node + flag) holding the source features */
  const int32_t* gdst;     /* [E] same for the destination features */
  const int32_t* vdst_ptr; /* [2N+1] CSR over gather rows of the sorted positions (by gdst) */
  const int32_t* vsrc_ptr; /* [2N+1] CSR over gather rows of vsrc_pos (by gsrc) */
  const int32_t* vsrc_pos; /* [E] sorted positions grouped by gsrc */
  int64_t num_gather_rows; /* 2N (autoregressive) or 0 */
} gcpnet_graph;

/* Sizes the caller must allocate for one layer on one graph (all in floats unless noted). */
typedef struct gcpnet_plan {
  int32_t edge_tile, edge_grid_fwd, edge_grid_bwd, node_tile, node_grid_fwd, node_grid_bwd;
  int32_t edge_smem_fwd_bytes, edge_smem_bwd_bytes, node_smem_fwd_bytes, node_smem_bwd_bytes;
  int64_t agg_floats;            /* per-destination message sums [N][s+3v] + two carry rows per edge tile (the per-edge
                                    messages are summed inside the edge kernel's tiles and never written) */
  int64_t saved_edge_floats;     /* activations kept for backward (0 in inference) */
  int64_t saved_node_floats;
  int64_t edge_partial_floats;   /* edge_grid_bwd * n_edge_params */
  int64_t node_partial_floats;   /* node_grid_bwd * n_node_params */
  int64_t edge_cotangent_floats; /* 2 * E * (s+3v): per-edge cotangents of the gathered node features */
  int64_t agg_cotangent_floats;  /* N * (s+3v) */
  int64_t packed_floats;         /* packed (chunked, padded) copy of the layer's weights, rewritten by every forward */
  int32_t tc_edge_path;          /* 1: this layer's edge kernels can run on the tcgen05 tensor-core path */
  int32_t reserved;
  int64_t prenorm_floats;        /* pre_norm: N * (s+3v) normalised layer input (forward, kept for backward) */
  int64_t prenorm_ws_floats;     /* pre_norm: backward workspace (cotangent of the normalised input, row statistics, partials) */
  int64_t edge_spill_floats;     /* FFMA edge backward: operand rows of the off-tile weight-gradient product (0: not used) */
} gcpnet_plan;

typedef struct gcpnet_forward_io {
  const float *h, *chi, *e, *xi, *frames, *pos;  /* pos may be NULL when !has_pos */
  float *out_h, *out_chi, *out_pos;
  float* agg;          /* plan.agg_floats: segment sums of the messages, written by the edge kernel */
  float* saved_edge;   /* plan.saved_edge_floats or NULL (inference) */
  float* saved_node;   /* plan.saved_node_floats or NULL (inference) */
  float* packed;       /* plan.packed_floats: written by the forward call, read by the matching backward */
  int32_t packed_ready; /* 1: `packed` already holds this layer's packed weights (gcpnet_layer_pack): skip the pack kernels */
  int32_t reserved;
  const float *h_gather, *chi_gather; /* autoregressive: [2N][s], [2N][v][3] gather table, row 2i = node_rep[i], row 2i+1 =
                                         node_rep_regressive[i] (gcpnet.py:1065-1116); NULL otherwise */
  float* prenorm;       /* plan.prenorm_floats (pre_norm layers) */
} gcpnet_forward_io;

typedef struct gcpnet_backward_io {
  const float *h, *chi, *e, *xi, *frames;        /* the forward inputs */
  const float *saved_edge, *saved_node;          /* written by gcpnet_layer_forward */
  const float *g_out_h, *g_out_chi, *g_out_pos;  /* cotangents of the outputs (g_out_pos NULL when !has_pos) */
  float *g_h, *g_chi, *g_e, *g_xi;               /* cotangents of the inputs */
  float* g_params;                               /* [n_edge_params + n_node_params] flat parameter gradient (overwritten) */
  float *ws_agg, *ws_edge, *ws_edge_partial, *ws_node_partial; /* workspaces sized by the plan */
  const float* packed;                           /* the forward call's packed weights */
  const float *h_gather, *chi_gather;            /* autoregressive: the forward's gather table */
  float *g_h_gather, *g_chi_gather;              /* autoregressive: [2N] cotangent of the gather table; rows 2i additionally
                                                    carry the direct cotangent of node i, g_h / g_chi are then scratch */
  const float* prenorm;                          /* pre_norm: the forward's normalised input */
  float* ws_prenorm;                             /* plan.prenorm_ws_floats */
  float* ws_edge_spill;                          /* plan.edge_spill_floats (FFMA edge kernels): per message GCP the rows of gT, Z and
                                                    gg of every edge; one output-parallel product over all edges then forms the
                                                    scalar_out / vector_out_scale weight gradients instead of a read-modify-write
                                                    of per-CTA partials per tile.  NULL: the tiles form them (slower) */
} gcpnet_backward_io;

int gcpnet_version(void);
const char* gcpnet_last_error(void);
/* number of kernels this library has enqueued so far in this process (bench.py: gpu_launches) */
uint64_t gcpnet_launch_count(void);
/* Optional per-kernel timing with CUDA events on the launching stream (bench.py roofline leg).
 * which: 0 edge forward, 1 node forward, 2 node backward, 3 edge backward, 4 cotangent reduce,
 * 5 partial reduce, 6 graph build.  read() waits for the recorded events, returns their summed
 * elapsed time and count, and clears them.  Keep disabled during CUDA-graph capture. */
void gcpnet_profile_enable(int on);
/* Runtime options: "tc" = 1/0 use / do not use the tensor-core (tcgen05, 3xTF32) edge kernels where the plan
 * allows them (default 1).  Returns the previous value, -1 for an unknown option. */
int gcpnet_set_option(const char* name, int value);
/* Optional side stream.  When set, gcpnet_layer_backward enqueues the work that only feeds the PARAMETER gradient
 * (node-level weight-gradient partials, reduction of the per-CTA partials, chain rule to the reference's tensors) on
 * this stream, forked from / ordered after the caller's stream by events, so it overlaps with the next layer's
 * backward.  The caller must call gcpnet_join(stream) -- `stream` then waits for everything forked so far -- before
 * reading g_params, and must keep the backward workspaces alive until then.  NULL (default): everything runs in
 * order on the caller's stream.  At most 32 backward calls may be pending between two joins. */
int gcpnet_set_side_stream(void* stream);
int gcpnet_join(void* stream);
/* Development aid: when non-NULL, CTA 0 of the tensor-core edge kernels writes clock64() stamps of its first tile's
 * stages into this device buffer of 24 x 16 int64 (scripts/tc_bwd_stamps.py decodes them).  NULL switches it off. */
void gcpnet_debug_stamps(long long* device_buffer);
int gcpnet_profile_read(int which, double* total_ms, int64_t* launches);

/* CSR build (replaces the index side of torch_scatter.scatter, gcpnet.py:946 and comp/__init__.py:316). */
size_t gcpnet_graph_workspace_bytes(int64_t num_edges, int64_t num_nodes);
int gcpnet_graph_build(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, const float* frames,
                       int32_t* perm, int32_t* src, int32_t* dst, int32_t* dst_ptr, int32_t* src_pos,
                       int32_t* src_ptr, float* fbar, void* workspace, size_t workspace_bytes, void* stream);

/* Autoregressive layers (GCPInteractions.autoregressive_forward, gcpnet.py:1065-1116): edges with row < col read node_rep
 * at both ends, the others node_rep_regressive; both message sets are summed per destination and divided by the in-degree
 * (clamped at 1).  Same outputs as gcpnet_graph_build (sorted by destination; inside a destination, row < col edges
 * first) plus the gather views of gcpnet_graph.  No boolean-mask indexing, no host synchronisation. */
size_t gcpnet_graph_ar_workspace_bytes(int64_t num_edges, int64_t num_nodes);
int gcpnet_graph_build_autoregressive(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, const float* frames,
                                      int32_t* perm, int32_t* src, int32_t* dst, int32_t* dst_ptr, int32_t* src_pos,
                                      int32_t* src_ptr, float* fbar, int32_t* gsrc, int32_t* gdst, int32_t* vdst_ptr,
                                      int32_t* vsrc_ptr, void* workspace, size_t workspace_bytes, void* stream);

/* Node mask (gcpnet.py:1202-1217; comp/__init__.py:294-300): from the views of gcpnet_graph_build and node_mask[N] derive
 *   frames_eff[E][9]  frames with the rows of edges that touch a masked-out node zeroed (predicated: masked frames may
 *                     hold inf, comp/__init__.py:232-236) -- what the message GCPs scalarise with,
 *   fbar_ff[N][9]     mean frame of the feed-forward GCPs on the mask's subgraph, relabelled nodes indexing the ORIGINAL
 *                     mask exactly as the reference does (gcpnet.py:1232-1239 passes node_mask to scalarize),
 *   fbar_pos[N][9]    mean frame of the position-update GCP (full graph, masked edges contribute zero frames).
 * workspace: (num_nodes + 1) * 4 bytes + gcpnet_graph_workspace_bytes. */
int gcpnet_graph_mask(const int64_t* edge_index, int64_t num_edges, int64_t num_nodes, const float* frames,
                      const uint8_t* node_mask, const gcpnet_graph* graph, float* frames_eff, float* fbar_ff,
                      float* fbar_pos, void* workspace, size_t workspace_bytes, void* stream);

/* frames = localize(x, edge_index) (comp/__init__.py:220-269); node_mask may be NULL; with a mask the frames of edges that
 * touch a masked-out node are +inf, as in the reference. */
int gcpnet_localize(const float* pos, const int64_t* edge_index, int64_t num_edges, int norm_x_diff,
                    float* frames, void* stream);
int gcpnet_localize_masked(const float* pos, const int64_t* edge_index, int64_t num_edges, int norm_x_diff,
                           const uint8_t* node_mask, float* frames, void* stream);

/* centralize / decentralize (comp/__init__.py:170-217): centroid[g] = mean of the (unmasked) rows of pos with
 * batch_index == g; centered = pos - centroid[batch] (masked-out rows: +inf).  batch_index must be non-decreasing (a PyG
 * Batch guarantees it).  decentralize adds the centroids back. */
int gcpnet_centralize(const float* pos, const int64_t* batch_index, int64_t num_nodes, int64_t num_graphs,
                      const uint8_t* node_mask, float* centroid, float* centered, void* stream);
int gcpnet_decentralize(const float* pos, const int64_t* batch_index, int64_t num_nodes, const float* centroid,
                        const uint8_t* node_mask, float* out, void* stream);

/* Pack this layer's weights into `packed` (plan.packed_floats) ahead of time -- the weights are known at the start of a
 * step, so a caller can run all layers' packing on a side stream while the first layers compute, then pass
 * packed_ready = 1 to gcpnet_layer_forward.  Must be repeated whenever the parameters change. */
int gcpnet_layer_pack(const gcpnet_layer* layer, const gcpnet_plan* plan, int64_t num_nodes, int64_t num_edges, float* packed,
                      void* stream);

/* GCPInteractions.forward / its backward (gcpnet.py:1160-1262). */
int gcpnet_layer_plan(const gcpnet_layer* layer, int64_t num_nodes, int64_t num_edges, gcpnet_plan* plan);
int gcpnet_layer_forward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                         const gcpnet_forward_io* io, void* stream);
int gcpnet_layer_backward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                          const gcpnet_backward_io* io, void* stream);

/* GCPMessagePassing.forward alone (message + aggregate, gcpnet.py:949-960): out[N][s+3v]. */
int gcpnet_message_passing_forward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                                   const gcpnet_forward_io* io, float* aggregate, void* stream);

/* One GCP2 on its own (GCP2.forward, gcpnet.py:393-468), e.g. the edge / node embeddings of GCPEmbedding (gcpnet.py:735-823).
 * Rows are entities: s_in[M][si], v_in[M][vi][3], frames[M][9] = the edge's frame (node_inputs=False) or the mean frame over
 * the node's outgoing edges (node_inputs=True: gcpnet_graph.fbar / fbar_pos).  op->grad_off[] are offsets into this module's
 * own flat gradient g_params[gcpnet_gcp2_plan.n_params]. */
typedef struct gcpnet_gcp2_plan {
  int32_t tile, grid, smem_fwd_bytes, smem_bwd_bytes, n_params, reserved;
  int64_t packed_floats;   /* packed weights (written by the forward call, read by the matching backward) */
  int64_t saved_floats;    /* M * (so + vo) activations kept for backward */
  int64_t partial_floats;  /* grid * n_params per-CTA weight-gradient partials */
} gcpnet_gcp2_plan;
int gcpnet_gcp2_plan_query(const gcpnet_gcp2* op, int64_t num_rows, gcpnet_gcp2_plan* plan);
int gcpnet_gcp2_forward(const gcpnet_gcp2* op, int64_t num_rows, const float* s_in, const float* v_in, const float* frames,
                        int enable_e3, float slope, float* s_out, float* v_out, float* saved, float* packed, void* stream);
int gcpnet_gcp2_backward(const gcpnet_gcp2* op, int64_t num_rows, const float* s_in, const float* v_in, const float* frames,
                         int enable_e3, float slope, const float* saved, const float* packed, const float* g_s_out,
                         const float* g_v_out, float* g_s_in, float* g_v_in, float* g_params, float* ws_partial, void* stream);

/* GCPLayerNorm on its own (comp/__init__.py:138-167): scalars nn.LayerNorm(s) (w, b, eps 1e-5), vectors divided by
 * sqrt(mean_c max(|v_c|^2, 1e-8)).  s == 0 or v == 0 drops that part.  Backward workspace: 2 N + 128 s floats. */
int gcpnet_layernorm_forward(const float* h, const float* chi, int64_t num_rows, int32_t s, int32_t v, const float* w,
                             const float* b, float* out_h, float* out_chi, void* stream);
int gcpnet_layernorm_backward(const float* h, const float* chi, int64_t num_rows, int32_t s, int32_t v, const float* w,
                              const float* g_out_h, const float* g_out_chi, float* g_h, float* g_chi, float* g_w, float* g_b,
                              float* workspace, void* stream);

/* Mean of the flat parameter gradient over the ranks of ONE node, over NVLink peer memory (csrc/p2p.cu): the path's only
 * exchange step (Lightning DDP / NCCL in the reference).  create: allocates this rank's IPC buffer
 * (num_floats gradient floats + flags), returns its device pointer (the layers' gradient sink) and its 64-byte CUDA IPC
 * handle; the caller exchanges the handles (any transport), then connect() maps the peers.  allreduce_mean enqueues one
 * kernel on `stream`: out[offset .. offset + count) = mean over ranks of their buffers' [offset .. offset + count).  Every
 * rank must issue the same sequence of calls.  No allocation or synchronisation per call: CUDA-graph capturable. */
int gcpnet_p2p_create(int rank, int world, int64_t num_floats, void** handle_out, void** data_out, unsigned char ipc_handle[64]);
int gcpnet_p2p_connect(void* handle, const unsigned char* all_handles, float* out);
int gcpnet_p2p_allreduce_mean(void* handle, int64_t offset, int64_t count, void* stream);
int gcpnet_p2p_destroy(void* handle);

/* Backward of GCPMessagePassing.forward alone: g_aggregate[N][s+3v] = cotangent of the aggregate.  Uses of `io`:
 * h, chi, e, xi, frames, saved_edge, packed (inputs); g_h, g_chi, g_e, g_xi (overwritten); g_params[0 .. n_edge_params)
 * (the message_fusion.* gradients); ws_edge, ws_edge_partial (workspaces).  The node-side fields are ignored. */
int gcpnet_message_passing_backward(const gcpnet_layer* layer, const gcpnet_graph* graph, const gcpnet_plan* plan,
                                    const gcpnet_backward_io* io, const float* g_aggregate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GCPNET_B200_H */
