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

// One B tile (or bias vector) of the packed blob.  kind 0: matrix with up to three column ranges
// (tile columns [tc0, tc0+len) <- reference columns [rc0, rc0+len)); kind 1: vector_down tile (rows < hd
// from vector_down.weight, rows 13..15 from vector_down_frames.weight); kind 2: vector (bias).
struct TcPackItem {
  int kind, dst_hi, dst_lo;
  int R, C;            // tile rows (slab pitch) / columns
  int nreal, ldw, hd;  // real rows, row stride of the source, (kind 1) hidden dim
  int nrange, tc0[3], rc0[3], len[3];
  const float *w, *w2;
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
  int max_sm = 0, max_w = 0, zcols = 0, tcols = 0;
  p.ring_s.n = 0; p.ring_w.n = 0;
  for (int k = 0; k < L; ++k) {
    const gcpnet_gcp2& d = l.message[k];
    TcGcp& g = p.g[k];
    g.si = d.si; g.vi = d.vi; g.so = d.so; g.vo = d.vo; g.hd = d.hd;
    g.act_s = d.act_s; g.act_v = d.act_v; g.vres = d.vector_residual;
    if (d.hd < 1 || d.hd > DCOL) return no("hidden vector width beyond 13");
    if (d.so != s || d.vo != v) return no("message GCP output dims differ from node dims");
    if (k == 0 && (d.si != 2 * s + se || d.vi != 2 * v + ve || d.vector_residual)) return no("unexpected message GCP 0 shape");
    if (k > 0 && (d.si != s || d.vi != v)) return no("unexpected message GCP shape");
    g.hdp = rup(d.hd, 8); g.sop = rup(d.so, 16); g.vop = rup(d.vo, 16); g.gk = rup(d.so, 8);
    const int Kref = d.si + d.hd + 9;
    // ---- small chunk
    const int sm0 = cursor;
    g.nvseg = k == 0 ? 3 : 1;
    const int vch0[3] = {0, v, v + ve};
    for (int i = 0; i < g.nvseg; ++i) {
      const int nch = k == 0 ? (i == 1 ? ve : v) : v;
      g.vkc[i] = rup(nch, 8);
      const int fl = 16 * g.vkc[i];
      const int ohi = take(fl), olo = take(fl);
      g.o_wd_hi[i] = ohi - sm0; g.o_wd_lo[i] = olo - sm0;
      TcPackItem& it = item();
      it = TcPackItem{1, ohi, olo, 16, g.vkc[i], 16, d.vi, d.hd, 1, {0, 0, 0}, {k == 0 ? vch0[i] : 0, 0, 0}, {nch, 0, 0}, d.vector_down, d.vector_down_frames};
    }
    {
      const int fl = g.vop * g.hdp;
      const int ohi = take(fl), olo = take(fl);
      g.o_wu_hi = ohi - sm0; g.o_wu_lo = olo - sm0;
      item() = TcPackItem{0, ohi, olo, g.vop, g.hdp, d.vo, d.hd, 0, 1, {0, 0, 0}, {0, 0, 0}, {d.hd, 0, 0}, d.vector_up, nullptr};
    }
    {
      const int fl = g.vop * g.gk;
      const int ohi = take(fl), olo = take(fl);
      g.o_wg_hi = ohi - sm0; g.o_wg_lo = olo - sm0;
      item() = TcPackItem{0, ohi, olo, g.vop, g.gk, d.vo, d.so, 0, 1, {0, 0, 0}, {0, 0, 0}, {d.so, 0, 0}, d.vector_out_scale_w, nullptr};
    }
    {
      const int obs = take(g.sop), obg = take(g.vop);
      g.o_bs = obs - sm0; g.o_bg = obg - sm0;
      item() = TcPackItem{2, obs, 0, g.sop, 1, d.so, 0, 0, 0, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, d.scalar_out_b, nullptr};
      item() = TcPackItem{2, obg, 0, g.vop, 1, d.vo, 0, 0, 0, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, d.vector_out_scale_b, nullptr};
    }
    const int smfl = cursor - sm0;
    if (p.ring_s.n >= MAX_RSEQ) return no("too many small chunks");
    p.ring_s.c[p.ring_s.n++] = TcChunk{sm0, smfl};
    if (smfl > max_sm) max_sm = smfl;
    // ---- scalar_out K-segments
    auto add_seg = [&](int a_tile, int kc, int nr, const int* tc0, const int* rc0, const int* len) {
      TcSeg& sgm = g.seg[g.nseg++];
      sgm.a_tile = a_tile; sgm.kc = kc;
      const int fl = g.sop * kc;
      const int ohi = take(fl), olo = take(fl);
      TcPackItem& it = item();
      it = TcPackItem{0, ohi, olo, g.sop, kc, d.so, Kref, 0, nr, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, d.scalar_out_w, nullptr};
      for (int i = 0; i < nr; ++i) { it.tc0[i] = tc0[i]; it.rc0[i] = rc0[i]; it.len[i] = len[i]; }
      p.ring_w.c[p.ring_w.n++] = TcChunk{ohi, fl};
      p.ring_w.c[p.ring_w.n++] = TcChunk{olo, fl};
      if (fl > max_w) max_w = fl;
    };
    g.nseg = 0;
    if (p.ring_w.n + 6 > MAX_RSEQ || P.pack.n + 4 > MAX_PACK_ITEMS) return no("too many weight chunks");
    if (k == 0) {
      // reference column order of scalar_out.weight: [h_row (s) | e (se) | h_col (s) | n (hd) | q (9)]   (gcpnet.py:917,422,436)
      g.zc0 = se; g.kz = rup(se + d.hd + 9, 8);
      const int tc0[2] = {0, se}, rc0[2] = {s, 2 * s + se}, len[2] = {se, d.hd + 9};
      add_seg(0, g.kz, 2, tc0, rc0, len);
      const int z = 0, r1 = 0, r2 = s + se, ls = s;
      add_seg(1, rup(s, 8), 1, &z, &r1, &ls);
      add_seg(2, rup(s, 8), 1, &z, &r2, &ls);
      if (rup(s, 8) > tcols) tcols = rup(s, 8);
      if (3 * PW > tcols) tcols = 3 * PW;   // chi_col planes are staged in the T tile
      if (3 * PW > zcols) zcols = 3 * PW;   // chi_row planes are staged in the Z tile
    } else {
      g.zc0 = d.si; g.kz = rup(Kref, 8);
      const int z = 0, ln = Kref;
      add_seg(0, g.kz, 1, &z, &z, &ln);
    }
    if (g.kz > zcols) zcols = g.kz;
    if (g.gk > tcols) tcols = g.gk;
    if (s > zcols) zcols = s;
  }
  P.pack.n = P.pack.n;
  P.blob_floats = cursor;
  // ---- rings
  p.ring_s.nslot = 2; p.ring_s.slot_floats = rup(max_sm, 32);
  p.ring_w.nslot = 3; p.ring_w.slot_floats = rup(max_w, 32);
  // ---- shared-memory map
  int off = 0;
  auto carve = [&](int floats) { const int o = off; off += rup(floats, 32); return o; };
  p.ZBUF = carve((zcols / 4) * SLAB);
  p.TBUF = carve((tcols / 4) * SLAB);
  p.VBUF = carve(3 * PLANE);
  p.HBUF = carve(3 * PLANE);
  p.FBUF = carve(TE * 9);
  p.RING_S = carve(p.ring_s.nslot * p.ring_s.slot_floats);
  p.RING_W = carve(p.ring_w.nslot * p.ring_w.slot_floats);
  p.BARS = carve(2 * (1 + p.ring_s.nslot + p.ring_w.nslot));
  p.smem_floats = off;
  if ((long long)off * 4 > TC_SMEM_LIMIT_BYTES) return no("tile does not fit shared memory");
  // ---- TMEM map (columns)
  int col = 0;
  auto tcol = [&](int n) { const int o = col; col += n; return o; };
  p.ZLO = tcol(zcols); p.TLO = tcol(tcols); p.VLO = tcol(3 * PW); p.HLO = tcol(3 * PW);
  p.HDACC = tcol(3 * 16); p.TACC = tcol(rup(s, 16)); p.GACC = tcol(16); p.UACC = tcol(3 * 16);
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
