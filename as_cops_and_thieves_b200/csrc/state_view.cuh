// state_view.cuh — SoA <-> packed-record conversion of the world state (parity tests, checkpoints).
// Included by cat_b200.cu inside its anonymous namespace.
#pragma once
// ------------------------------------------------------------------ state pack / unpack
struct ViewParams {
  float* state;
  int rec_words, n_worlds, A, P;
  int o_vel, o_vb, o_tc, o_wkey, o_wjn, o_page, o_pjn, o_sc, o_ep, o_flags;
  CatStateView v;
  int set;
};

__global__ void cat_state_view_kernel(const ViewParams p) {
  const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= p.n_worlds) return;
  float* rec = p.state + (size_t)w * p.rec_words;
  uint32_t* recu = reinterpret_cast<uint32_t*>(rec);
  const int A = p.A, P = p.P;
  auto xfer = [&](float* ext, int off, int n) {
    if (!ext) return;
    for (int i = 0; i < n; ++i) { if (p.set) rec[off + i] = ext[(size_t)w * n + i]; else ext[(size_t)w * n + i] = rec[off + i]; }
  };
  xfer(p.v.pos, 0, 2 * A);
  if (p.set && p.v.pos) recu[p.o_flags] = 0xFFu;  // positions changed: re-evaluate the alpha = 0 candidates
  xfer(p.v.vel, p.o_vel, 2 * A);
  xfer(p.v.vbias, p.o_vb, 2 * A);
  xfer(p.v.tc, p.o_tc, 2 * A);
  xfer(p.v.wall_jn, p.o_wjn, A * kSlots);
  xfer(p.v.pair_jn, p.o_pjn, P);
  if (p.v.step_count) { if (p.set) recu[p.o_sc] = (uint32_t)p.v.step_count[w]; else p.v.step_count[w] = (int32_t)recu[p.o_sc]; }
  if (p.v.episode) { if (p.set) recu[p.o_ep] = p.v.episode[w]; else p.v.episode[w] = recu[p.o_ep]; }
  if (p.v.wall_hull && p.v.wall_age) {
    for (int i = 0; i < A * kSlots; ++i) {
      const size_t e = (size_t)w * A * kSlots + i;
      if (p.set) {
        recu[p.o_wkey + i] = p.v.wall_hull[e] < 0 ? kEmpty : ((uint32_t)p.v.wall_hull[e] & 0xFFFF) | ((uint32_t)p.v.wall_age[e] << 16);
      } else {
        const uint32_t key = recu[p.o_wkey + i];
        p.v.wall_hull[e] = key == kEmpty ? -1 : (int32_t)(key & 0xFFFF);
        p.v.wall_age[e] = key == kEmpty ? -1 : (int32_t)(key >> 16);
      }
    }
  }
  if (p.v.pair_age) {
    for (int i = 0; i < P; ++i) {
      const size_t e = (size_t)w * P + i;
      if (p.set) recu[p.o_page + i] = p.v.pair_age[e] < 0 ? kEmpty : (uint32_t)p.v.pair_age[e];
      else p.v.pair_age[e] = recu[p.o_page + i] == kEmpty ? -1 : (int32_t)recu[p.o_page + i];
    }
  }
}

