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
//   PK_SCALAR  W[R = sop + 16][C]: rows < so scalar_out.weight (w, row stride ldw); rows sop.. = vector_out_scale.weight (w2) . w
//   PK_VECTOR  W[R = 32][C]:       rows < hd vector_down.weight (w, ldw = vi); rows 13..15 vector_down_frames.weight (w2);
//                                  rows 16.. = vector_up.weight (w3) . w
//   PK_BIAS    [sop + 16]:         scalar_out.bias (w), then vector_out_scale.bias (w3) + w2 . w
// tr = 1 stores the transpose (tile rows = Rt >= C columns of W, tile columns = R): the B operand of the data-gradient GEMM.
enum { PK_SCALAR = 0, PK_VECTOR = 1, PK_BIAS = 2 };
struct TcPackItem {
  int kind, dst_hi, dst_lo;       // dst_lo < 0: no lo part
  int R, C, tr, Rt;
  int nreal, ldw, hd, vo;         // so (PK_SCALAR, PK_BIAS); row stride of w; hidden dim; vector outputs
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
  long long pq_floats = 0;    // N * (2 * pw + 192)
  // backward
  TcBwdParams bproto{};
  long long y_floats = 0;       // tiles * (y_img_g + y_img_v): per-edge cotangents of message GCP 0's per-node products
  long long partial_floats = 0; // grid * partial_stride
  long long a_floats = 0;       // N * 2 * (pw + 96): per-node sums of Y over outgoing / incoming edges
  int node_partial_ctas = 0, node_partial_stride = 0;  // node-level weight-gradient partials of message GCP 0
  int wt_off_hi[12] = {0}, wt_off_lo[12] = {0};
};
constexpr int TC_POST_CTAS = 64;  // node chunks of the node-level weight-gradient partials

inline int rup(int x, int m) { return (x + m - 1) / m * m; }

inline TcPlan make_tc_plan(const gcpnet_layer& l, long long N, long long E) {
  TcPlan P;
  auto no = [&](const char* m) { P.why = m; return P; };
  const int L = l.num_message_layers, s = l.s, v = l.v, se = l.se, ve = l.ve;
  if (s % 16 || se % 8 || v % 4 || ve % 4) return no("dims not multiples of (16, 8, 4, 4)");
  if (v > PW || ve > 8) return no("vector channels beyond 16 / 8");
  if (L < 1) return no("no message layers");
  if (l.autoregressive) return no("autoregressive layers gather from two node tables (FFMA edge kernels)");
  if (l.attn_w != nullptr) return no("scalar message attention runs in the FFMA edge kernels");
  TcEdgeParams& p = P.proto;
  p.L = L; p.s = s; p.v = v; p.se = se; p.ve = ve;
  p.residual = l.residual_messages; p.e3 = l.enable_e3; p.slope = l.slope;
  p.pw = s + 16;
  int cursor = 0;
  auto take = [&](int floats) { const int o = cursor; cursor += rup(floats, 32); return o; };
  auto item = [&]() -> TcPackItem& { return P.pack.it[P.pack.n++]; };
  int max_sm = 0, max_w = 0, zcols = s;
  p.ring_s.n = 0; p.ring_w.n = 0;
  for (int k = 0; k < L; ++k) {
    const gcpnet_gcp2& d = l.message[k];
    TcGcp& g = p.g[k];
    g.si = d.si; g.vi = d.vi; g.so = d.so; g.vo = d.vo; g.hd = d.hd;
    g.act_s = d.act_s; g.vres = d.vector_residual;
    if (d.act_v != 0) return no("vector nonlinearity is not the identity (gate not composable)");
    if (d.flags != 0) return no("GCP-Baseline variants (no frame scalars / no vector gate) run the FFMA edge kernels");
    if (d.hd < 1 || d.hd > NSLOT) return no("hidden vector width beyond 12");
    if (d.so != s || d.vo != v) return no("message GCP output dims differ from node dims");
    if (k == 0 && (d.si != 2 * s + se || d.vi != 2 * v + ve || d.vector_residual)) return no("unexpected message GCP 0 shape");
    if (k > 0 && (d.si != s || d.vi != v)) return no("unexpected message GCP shape");
    g.sop = s;
    g.nslot = rup(d.hd, 4);
    g.zc0 = k == 0 ? se : s;
    g.kz = g.zc0 + g.nslot + 12;
    g.vkc = rup(k == 0 ? ve : v, 8);
    const int Kref = d.si + d.hd + 9, R = s + 16;
    if (p.ring_w.n + 2 > MAX_RSEQ || P.pack.n + 10 > MAX_PACK_ITEMS || p.ring_s.n >= MAX_RSEQ) return no("too many weight chunks");
    // ---- small chunk: vector batch tiles (forward and transposed), composed bias
    const int sm0 = cursor;
    {
      const int ch0 = k == 0 ? v : 0, nch = k == 0 ? ve : v;
      const int fl = VN * g.vkc, rt = rup(g.vkc, 16), flt = rt * VN;
      const int ohi = take(fl), olo = take(fl), thi = take(flt), tlo = take(flt);
      g.o_wv_hi = ohi - sm0; g.o_wv_lo = olo - sm0; g.o_wvt_hi = thi - sm0; g.o_wvt_lo = tlo - sm0;
      item() = TcPackItem{PK_VECTOR, ohi, olo, VN, g.vkc, 0, 0, 0, d.vi, d.hd, d.vo, 1, {0, 0, 0}, {ch0, 0, 0}, {nch, 0, 0},
                          d.vector_down, d.vector_down_frames, d.vector_up};
      item() = TcPackItem{PK_VECTOR, thi, tlo, VN, g.vkc, 1, rt, 0, d.vi, d.hd, d.vo, 1, {0, 0, 0}, {ch0, 0, 0}, {nch, 0, 0},
                          d.vector_down, d.vector_down_frames, d.vector_up};
      const int ob = take(R);
      g.o_b = ob - sm0;
      item() = TcPackItem{PK_BIAS, ob, -1, s, 1, 0, 0, d.so, 0, 0, d.vo, 0, {0, 0, 0}, {0, 0, 0}, {0, 0, 0},
                          d.scalar_out_b, d.vector_out_scale_w, d.vector_out_scale_b};
    }
    const int smfl = cursor - sm0;
    p.ring_s.c[p.ring_s.n++] = TcChunk{sm0, smfl};
    if (smfl > max_sm) max_sm = smfl;
    // ---- scalar batch tile over the Z-tile columns.  Reference column order of scalar_out.weight: [scalars (si) | n (hd) |
    // q (9)], for GCP 0 the scalars are [h_row (s) | e (se) | h_col (s)]  (gcpnet.py:917,422,436); tile tail: nslot | 9 | 3 zeros
    const int tc0[3] = {0, g.zc0, g.zc0 + g.nslot};
    const int rc0[3] = {k == 0 ? s : 0, d.si, d.si + d.hd};
    const int len[3] = {k == 0 ? se : s, d.hd, 9};
    {
      const int fl = R * g.kz;
      const int ohi = take(fl), olo = take(fl);
      TcPackItem& it = item();
      it = TcPackItem{PK_SCALAR, ohi, olo, R, g.kz, 0, 0, d.so, Kref, 0, d.vo, 3, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, d.scalar_out_w, d.vector_out_scale_w, nullptr};
      for (int i = 0; i < 3; ++i) { it.tc0[i] = tc0[i]; it.rc0[i] = rc0[i]; it.len[i] = len[i]; }
      p.ring_w.c[p.ring_w.n++] = TcChunk{ohi, fl};
      p.ring_w.c[p.ring_w.n++] = TcChunk{olo, fl};
      if (fl > max_w) max_w = fl;
      // transposed tile [kz -> 16][pw] for the data-gradient GEMM
      const int rt = rup(g.kz, 16), flt = rt * R;
      const int thi = take(flt), tlo = take(flt);
      TcPackItem& tt = item();
      tt = it;
      tt.dst_hi = thi; tt.dst_lo = tlo; tt.tr = 1; tt.Rt = rt;
      P.wt_off_hi[k] = thi; P.wt_off_lo[k] = tlo;
      P.bproto.kzn[k] = rt;
      if (flt > max_w) max_w = flt;
    }
    if (k == 0) {
      // node-level tiles of GCP 0 (hi parts only): h_row / h_col column blocks, chi_row / chi_col channel blocks
      const int z = 0, c_src = 0, c_dst = s + se, ls = s;
      const int f1 = R * s, f2 = VN * rup(v, 8);
      p.nt.ps = take(f1); p.nt.pd = take(f1); p.nt.qs = take(f2); p.nt.qd = take(f2);
      TcPackItem a = TcPackItem{PK_SCALAR, p.nt.ps, -1, R, s, 0, 0, d.so, Kref, 0, d.vo, 1, {z, 0, 0}, {c_src, 0, 0}, {ls, 0, 0}, d.scalar_out_w, d.vector_out_scale_w, nullptr};
      item() = a;
      a.dst_hi = p.nt.pd; a.rc0[0] = c_dst;
      item() = a;
      TcPackItem b = TcPackItem{PK_VECTOR, p.nt.qs, -1, VN, rup(v, 8), 0, 0, 0, d.vi, d.hd, d.vo, 1, {0, 0, 0}, {0, 0, 0}, {v, 0, 0},
                                d.vector_down, d.vector_down_frames, d.vector_up};
      item() = b;
      b.dst_hi = p.nt.qd; b.rc0[0] = v + ve;
      item() = b;
    }
    if (g.kz > zcols) zcols = g.kz;
  }
  P.blob_floats = cursor;
  // ---- rings
  p.ring_s.nslot = 2; p.ring_s.slot_floats = rup(max_sm, 32);
  p.ring_w.nslot = 4; p.ring_w.slot_floats = rup(max_w, 32);
  // ---- shared-memory map
  int off = 0;
  auto carve = [&](int floats) { const int o = off; off += rup(floats, 32); return o; };
  p.ZBUF = carve((zcols / 4) * SLAB);
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
  p.ZLO = tcol(zcols); p.VLO = tcol(3 * PW); p.VACC = tcol(3 * VN); p.TACC = tcol(s + 16);
  if (col > 512) return no("tile does not fit tensor memory");
  p.tmem_cols = col <= 32 ? 32 : (col <= 64 ? 64 : (col <= 128 ? 128 : (col <= 256 ? 256 : 512)));
  // ---- saved activations: per tile, per GCP k < L-1: S image (s/4 slabs) + V image (3 planes)
  // (sizes depend on the tile height: set below, once `rows` is known)
  // tile height: at most 128 rows, chosen so that the persistent CTAs (one per SM) all run the same number of tiles: w =
  // waves at full height, rows = E / (148 w) rounded up to 8 (>= 32).  The per-tile latency chain does not depend on the
  // row count; the row-proportional stages (weight-gradient products, TMEM / shared-memory traffic) shrink with it.
  int rows = TE;
  if (E > 0) {
    const long long t128 = (E + TE - 1) / TE, w = (t128 + 147) / 148;
    rows = rup((int)((E + 148 * w - 1) / (148 * w)), 8);
    if (rows < 32) rows = 32;
    if (rows > TE) rows = TE;
  }
  p.rows = rows;
  // compact images: only the tile's `rows` rows of every 4-column slab are stored (image_store / image_load)
  p.s_img = (s / 4) * rows * 4; p.v_img = 3 * (PW / 4) * rows * 4;
  p.saved_tile_stride = (long long)(L - 1) * (p.s_img + p.v_img);
  const long long tiles = (E + rows - 1) / rows;
  P.saved_floats = tiles * p.saved_tile_stride;
  P.pq_floats = N * (2 * p.pw + 192);
  P.grid = (int)(tiles < 1 ? 1 : (tiles > 148 ? 148 : tiles));
  // ======================== backward ========================
  {
    TcBwdParams& b = P.bproto;
    b.f = p;
    TcEdgeParams& f = b.f;
    // rings in reverse GCP order; per GCP: forward tile (recompute) then transposed tile (data gradient)
    f.ring_s.n = 0; f.ring_w.n = 0;
    for (int k = L - 1; k >= 0; --k) {
      f.ring_s.c[f.ring_s.n++] = p.ring_s.c[k];
      const int R = s + 16;
      f.ring_w.c[f.ring_w.n++] = p.ring_w.c[2 * k];
      f.ring_w.c[f.ring_w.n++] = p.ring_w.c[2 * k + 1];
      f.ring_w.c[f.ring_w.n++] = TcChunk{P.wt_off_hi[k], b.kzn[k] * R};
      f.ring_w.c[f.ring_w.n++] = TcChunk{P.wt_off_lo[k], b.kzn[k] * R};
    }
    f.ring_w.nslot = 2;
    // shared memory
    int o2 = 0;
    auto carve2 = [&](int floats) { const int o = o2; o2 += rup(floats, 32); return o; };
    const int zc = zcols > p.pw ? zcols : p.pw;
    f.ZBUF = carve2((zc / 4) * SLAB);
    f.VBUF = carve2(3 * PLANE);
    b.GTG = carve2((p.pw / 4) * SLAB);
    b.GHDU = carve2(3 * (VN / 4) * SLAB);
    f.FBUF = carve2(TE * 9);
    f.RING_S = carve2(f.ring_s.nslot * f.ring_s.slot_floats);
    f.RING_W = carve2(f.ring_w.nslot * f.ring_w.slot_floats);
    f.BARS = carve2(2 * (3 + f.ring_s.nslot + f.ring_w.nslot));
    f.smem_floats = o2;
    if ((long long)o2 * 4 > TC_SMEM_LIMIT_BYTES) return no("backward tile does not fit shared memory");
    // TMEM: GS | GV | R1 = Z lo / [gT|gg] lo | R2 = V lo / gV accumulator | vector accumulator | R3 = [T|g] / gZ accumulator | [gH|gD|gU] lo
    int c2 = 0;
    auto tc2 = [&](int n) { const int o = c2; c2 += n; return o; };
    int r3 = p.pw;
    for (int k = 0; k < L; ++k) if (b.kzn[k] > r3) r3 = b.kzn[k];
    b.GS = tc2(s); b.GV = tc2(3 * PW);
    f.ZLO = tc2(zc); f.VLO = tc2(3 * PW); f.VACC = tc2(3 * VN); f.TACC = tc2(r3); b.GHDULO = tc2(3 * VN);
    if (c2 > 512) return no("backward tile does not fit tensor memory");
    f.tmem_cols = 512;
    // partial rows
    int po = 0;
    for (int k = 0; k < L; ++k) { b.off_tg[k] = po; po += p.pw * p.g[k].kz; b.off_v[k] = po; po += VN * 16; }
    b.partial_stride = rup(po, 32);
    b.y_img_g = (p.pw / 4) * rows * 4; b.y_img_v = 3 * (VN / 4) * rows * 4;
    b.reduce_mean = l.reduce_mean;
    P.y_floats = tiles * (long long)(b.y_img_g + b.y_img_v);
    P.partial_floats = (long long)P.grid * b.partial_stride;
    P.a_floats = N * 2LL * (p.pw + 96);
    P.node_partial_ctas = TC_POST_CTAS;
    P.node_partial_stride = 2 * (p.pw * s + VN * 16);
    if (l.enable_e3) return no("e3 frames not covered by the tensor-core backward");
  }
  P.ok = true;
  return P;
}

}  // namespace tc
}  // namespace gcp
