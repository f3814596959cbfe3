// tc_setup.h -- host-side planning for the tensor-core edge kernels (tc_edge.cuh): which layers they
// cover, the packed-weight layout (B tiles in slab layout, hi / lo split for 3xTF32), the chunk
// sequences of the two shared-memory rings, the shared-memory and TMEM maps.
#pragma once
#include <string>

#include "../../include/gcpnet_b200.h"
#include "tc_edge.cuh"

namespace gcp {
namespace tc {

constexpr int TC_SMEM_LIMIT_BYTES = 227 * 1024;
constexpr int MAX_PACK_ITEMS = 128;

// One B tile (or bias vector) of the packed blob, written by tc_pack_kernel (tc_edge_dev.cuh).  Matrix tiles take up
// to three column ranges: tile columns [tc0, tc0+len) <- reference columns [rc0, rc0+len).
//   PK_SCALAR  [sop + 16][C]: rows < so scalar_out.weight (w, row stride ldw); rows sop.. = vector_out_scale.weight (w2) . w
//   PK_VECTOR  [32][C]:       rows < hd vector_down.weight (w, ldw = vi); rows 13..15 vector_down_frames.weight (w2);
//                             rows 16.. = vector_up.weight (w3) . w
//   PK_BIAS    [sop + 16]:    scalar_out.bias (w), then vector_out_scale.bias (w3) + w2 . w
enum { PK_SCALAR = 0, PK_VECTOR = 1, PK_BIAS = 2 };
struct TcPackItem {
  int kind, dst_hi, dst_lo;
  int R, C;                // tile rows (slab pitch) / columns
  int nreal, ldw, hd, vo;  // so (PK_SCALAR, PK_BIAS); row stride of w; hidden dim; vector outputs
  int nrange, tc0[3], rc0[3], len[3];
  const float *w, *w2, *w3;
};
struct TcPackProg { int n; float* blob; TcPackItem it[MAX_PACK_ITEMS]; };

struct TcPlan {
  bool ok = false;
  std::string why;            // why not, when !ok
  TcEdgeParams proto{};       // everything except pointers, E-dependent fields and blob base
  TcPackProg pack{};          // dst offsets relative to the TC blob; w pointers filled
  int blob_floats = 0;
  int grid = 0;
  long long saved_floats = 0; // training: tiles * (L-1) * (s_img + v_img)
};

inline int rup(int x, int m) { return (x + m - 1) / m * m; }

inline TcPlan make_tc_plan(const gcpnet_layer& l, long long E) {
  TcPlan P;
  auto no = [&](const char* m) { P.why = m; return P; };
  const int L = l.num_message_layers, s = l.s, v = l.v, se = l.se, ve = l.ve;
  if (s % 8 || se % 8 || v % 4 || ve % 4) return no("dims not multiples of (8, 8, 4, 4)");
  if (v > PW || ve > 8) return no("vector channels beyond 16 / 8");
  if (L < 1) return no("no message layers");
  TcEdgeParams& p = P.proto;
  p.L = L; p.s = s; p.v = v; p.se = se; p.ve = ve;
  p.residual = l.residual_messages; p.e3 = l.enable_e3; p.slope = l.slope;
  int cursor = 0;
  auto take = [&](int floats) { const int o = cursor; cursor += rup(floats, 32); return o; };
  auto item = [&]() -> TcPackItem& { return P.pack.it[P.pack.n++]; };
  int max_sm = 0, max_w = 0, zcols = 3 * PW, xcols = 3 * PW;  // chi_row / chi_col planes are staged in the Z / X tiles
  p.ring_s.n = 0; p.ring_w.n = 0;
  for (int k = 0; k < L; ++k) {
    const gcpnet_gcp2& d = l.message[k];
    TcGcp& g = p.g[k];
    g.si = d.si; g.vi = d.vi; g.so = d.so; g.vo = d.vo; g.hd = d.hd;
    g.act_s = d.act_s; g.vres = d.vector_residual;
    if (d.act_v != 0) return no("vector nonlinearity is not the identity (gate not composable)");
    if (d.hd < 1 || d.hd > NSLOT) return no("hidden vector width beyond 12");
    if (d.so != s || d.vo != v) return no("message GCP output dims differ from node dims");
    if (k == 0 && (d.si != 2 * s + se || d.vi != 2 * v + ve || d.vector_residual)) return no("unexpected message GCP 0 shape");
    if (k > 0 && (d.si != s || d.vi != v)) return no("unexpected message GCP shape");
    g.sop = rup(d.so, 16);
    g.nslot = rup(d.hd, 4);
    const int ztail = g.nslot + 12;
    const int Kref = d.si + d.hd + 9;
    if (p.ring_w.n + 6 > MAX_RSEQ || P.pack.n + 8 > MAX_PACK_ITEMS || p.ring_s.n >= MAX_RSEQ) return no("too many weight chunks");
    // ---- small chunk: vector batch tiles, biases
    const int sm0 = cursor;
    g.nvseg = k == 0 ? 3 : 1;
    const int vch0[3] = {0, v, v + ve};
    for (int i = 0; i < g.nvseg; ++i) {
      const int nch = k == 0 ? (i == 1 ? ve : v) : v;
      g.vkc[i] = rup(nch, 8);
      const int fl = VN * g.vkc[i];
      const int ohi = take(fl), olo = take(fl);
      g.o_wd_hi[i] = ohi - sm0; g.o_wd_lo[i] = olo - sm0;
      item() = TcPackItem{PK_VECTOR, ohi, olo, VN, g.vkc[i], 0, d.vi, d.hd, d.vo, 1, {0, 0, 0}, {k == 0 ? vch0[i] : 0, 0, 0}, {nch, 0, 0},
                          d.vector_down, d.vector_down_frames, d.vector_up};
    }
    {
      const int ob = take(g.sop + 16);
      g.o_bs = ob - sm0; g.o_bg = ob - sm0 + g.sop;
      item() = TcPackItem{PK_BIAS, ob, 0, g.sop, 1, d.so, 0, 0, d.vo, 0, {0, 0, 0}, {0, 0, 0}, {0, 0, 0},
                          d.scalar_out_b, d.vector_out_scale_w, d.vector_out_scale_b};
    }
    const int smfl = cursor - sm0;
    p.ring_s.c[p.ring_s.n++] = TcChunk{sm0, smfl};
    if (smfl > max_sm) max_sm = smfl;
    // ---- scalar batch K-segments
    g.nseg = 0;
    auto add_seg = [&](int a_tile, int kc, int nr, const int* tc0, const int* rc0, const int* len) {
      TcSeg& sgm = g.seg[g.nseg++];
      sgm.a_tile = a_tile; sgm.kc = kc;
      const int R = g.sop + 16, fl = R * kc;
      const int ohi = take(fl), olo = take(fl);
      TcPackItem& it = item();
      it = TcPackItem{PK_SCALAR, ohi, olo, R, kc, d.so, Kref, 0, d.vo, nr, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, d.scalar_out_w, d.vector_out_scale_w, nullptr};
      for (int i = 0; i < nr; ++i) { it.tc0[i] = tc0[i]; it.rc0[i] = rc0[i]; it.len[i] = len[i]; }
      p.ring_w.c[p.ring_w.n++] = TcChunk{ohi, fl};
      p.ring_w.c[p.ring_w.n++] = TcChunk{olo, fl};
      if (fl > max_w) max_w = fl;
    };
    // reference column order of scalar_out.weight: [scalars (si) | n (hd) | q (9)], and for GCP 0 the scalars are
    // [h_row (s) | e (se) | h_col (s)]  (gcpnet.py:917,422,436); tile tail: hd -> 4 norm slots | 9 frame scalars | 3 zeros
    if (k == 0) {
      g.zc0 = se;
      const int tc0[3] = {0, se, se + g.nslot}, rc0[3] = {s, d.si, d.si + d.hd}, len[3] = {se, d.hd, 9};
      add_seg(0, se + ztail, 3, tc0, rc0, len);
      const int z = 0, r1 = 0, r2 = s + se, ls = s;
      add_seg(1, s, 1, &z, &r1, &ls);
      add_seg(2, s, 1, &z, &r2, &ls);
      if (s > xcols) xcols = s;
      if (se + ztail > zcols) zcols = se + ztail;
    } else {
      g.zc0 = d.si;
      const int tc0[3] = {0, d.si, d.si + g.nslot}, rc0[3] = {0, d.si, d.si + d.hd}, len[3] = {d.si, d.hd, 9};
      add_seg(0, d.si + ztail, 3, tc0, rc0, len);
    }
    if (s + ztail > zcols) zcols = s + ztail;
  }
  P.blob_floats = cursor;
  // ---- rings
  p.ring_s.nslot = 2; p.ring_s.slot_floats = rup(max_sm, 32);
  p.ring_w.nslot = 3; p.ring_w.slot_floats = rup(max_w, 32);
  // ---- shared-memory map
  int off = 0;
  auto carve = [&](int floats) { const int o = off; off += rup(floats, 32); return o; };
  p.ZBUF = carve((zcols / 4) * SLAB);
  p.XBUF = carve((xcols / 4) * SLAB);
  p.VBUF = carve(3 * PLANE);
  p.FBUF = carve(TE * 9);
  p.RING_S = carve(p.ring_s.nslot * p.ring_s.slot_floats);
  p.RING_W = carve(p.ring_w.nslot * p.ring_w.slot_floats);
  p.BARS = carve(2 * (1 + p.ring_s.nslot + p.ring_w.nslot));
  p.smem_floats = off;
  if ((long long)off * 4 > TC_SMEM_LIMIT_BYTES) return no("tile does not fit shared memory");
  // ---- TMEM map (columns)
  int col = 0;
  auto tcol = [&](int n) { const int o = col; col += n; return o; };
  p.ZLO = tcol(zcols); p.XLO = tcol(xcols); p.VLO = tcol(3 * PW);
  p.VACC = tcol(3 * VN); p.TACC = tcol(rup(s, 16) + 16);
  if (col > 512) return no("tile does not fit tensor memory");
  p.tmem_cols = col <= 32 ? 32 : (col <= 64 ? 64 : (col <= 128 ? 128 : (col <= 256 ? 256 : 512)));
  // ---- saved activations: per tile, per GCP k < L-1: S image (s/4 slabs) + V image (3 planes)
  p.s_img = (s / 4) * SLAB; p.v_img = 3 * PLANE;
  p.saved_tile_stride = (long long)(L - 1) * (p.s_img + p.v_img);
  const long long tiles = (E + TE - 1) / TE;
  P.saved_floats = tiles * p.saved_tile_stride;
  P.grid = (int)(tiles < 1 ? 1 : (tiles > 148 ? 148 : tiles));
  P.ok = true;
  return P;
}

}  // namespace tc
}  // namespace gcp
