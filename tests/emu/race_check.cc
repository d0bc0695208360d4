// ThreadSanitizer pass over the emulated kernels (TEST INFRASTRUCTURE ONLY): one CUDA thread = one std::thread and
// __syncthreads = std::barrier, so a missing barrier or an unsynchronised shared-memory reuse in a kernel shows up as a
// data race.  Built and run by tests/test_emu_kernels.py::test_emulated_kernels_are_race_free_under_tsan:
//   g++ -std=c++20 -O1 -g -fsanitize=thread -DIAF_EMU -I tests/emu -I iaf_b200/csrc -x c++ <sources> race_check.cc
// Runs weight packing, the forward SIMT step (two row bands), the training forward and every backward kernel once on a
// small Theano-variant stack (pad channel, flip) and once on a TF one.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/iaf_b200.h"

static std::vector<float> rnd(size_t n, float s, unsigned seed) {
  std::vector<float> v(n);
  unsigned x = seed * 2654435761u + 12345u;
  for (auto& e : v) {
    x = x * 1664525u + 1013904223u;
    e = s * ((float)((x >> 8) & 0xFFFF) / 32768.0f - 1.0f);
  }
  return v;
}

static int run(int variant, int n_z, int nh, int H, int W, int B) {
  iaf_desc_t d = {};
  d.variant = variant; d.n_z = n_z; d.n_hidden = 1; d.hidden[0] = nh; d.n_heads = 2; d.head[0] = d.head[1] = n_z;
  d.H = H; d.W = W; d.nl = IAF_NL_ELU; d.path = IAF_PATH_SIMT;
  iaf_plan_t* pl = nullptr;
  if (iaf_plan_create(&pl, &d) != IAF_OK) return 1;
  const int cin[3] = {n_z, nh, nh}, cout[3] = {nh, n_z, n_z};
  std::vector<std::vector<float>> w(3), s(3), b(3), gw(3), gs(3), gb(3);
  const float* wp[3]; const float* sp[3]; const float* bp[3];
  float* gwp[3]; float* gsp[3]; float* gbp[3];
  for (int i = 0; i < 3; ++i) {
    const size_t nw = variant == IAF_VARIANT_TF ? (size_t)9 * cin[i] * cout[i] : (size_t)cout[i] * (cin[i] + 1) * 9;
    w[i] = rnd(nw, 0.05f, 10 + i); s[i] = rnd(cout[i], 0.3f, 20 + i); b[i] = rnd(cout[i], 0.1f, 30 + i);
    gw[i].assign(nw, 0.f); gs[i].assign(cout[i], 0.f); gb[i].assign(cout[i], 0.f);
    wp[i] = w[i].data(); sp[i] = s[i].data(); bp[i] = b[i].data();
    gwp[i] = gw[i].data(); gsp[i] = gs[i].data(); gbp[i] = gb[i].data();
  }
  if (iaf_pack_weights(pl, wp, sp, bp, nullptr) != IAF_OK) return 2;
  const size_t nzv = (size_t)B * n_z * H * W, ncv = (size_t)B * nh * H * W;
  auto z = rnd(nzv, 1.f, 1), ctx = rnd(ncv, 0.1f, 2), g1 = rnd(nzv, 1.f, 3), g2 = rnd(nzv, 1.f, 4), g3 = rnd(B, 1.f, 5);
  std::vector<float> zo(nzv), ls(nzv), ld(B), hid(ncv), gz(nzv), gc(ncv);
  float* hp[1] = {hid.data()};
  const float* hcp[1] = {hid.data()};
  if (iaf_step_fwd(pl, z.data(), ctx.data(), zo.data(), ls.data(), ld.data(), B, nullptr) != IAF_OK) return 3;
  if (iaf_step_fwd_train(pl, z.data(), ctx.data(), zo.data(), ls.data(), ld.data(), hp, B, nullptr) != IAF_OK) return 4;
  if (iaf_step_bwd(pl, z.data(), ctx.data(), wp, sp, g1.data(), g2.data(), g3.data(), gz.data(), gc.data(), gwp, gsp, gbp, B,
                   nullptr) != IAF_OK) return 5;
  if (iaf_step_bwd_saved(pl, z.data(), zo.data(), ls.data(), hcp, wp, sp, g1.data(), g2.data(), g3.data(), gz.data(), gc.data(),
                         gwp, gsp, gbp, B, nullptr) != IAF_OK) return 6;
  std::vector<float> o0(nzv), o1(nzv), o2(nzv), o3(nzv), o4(nzv), klbc((size_t)B * n_z), klc(B);
  if (iaf_layer_fwd(pl, z.data(), g1.data(), ls.data(), g2.data(), ls.data(), ctx.data(), zo.data(), o0.data(), klbc.data(),
                    klc.data(), B, nullptr) != IAF_OK) return 7;
  if (iaf_layer_bwd(pl, z.data(), g1.data(), ls.data(), g2.data(), ls.data(), ctx.data(), wp, sp, g1.data(), g2.data(),
                    klbc.data(), klc.data(), o0.data(), o1.data(), o2.data(), o3.data(), o4.data(), gc.data(), gwp, gsp, gbp, B,
                    nullptr) != IAF_OK) return 8;
  iaf_plan_destroy(pl);
  return 0;
}

int main() {
  int rc = run(IAF_VARIANT_THEANO, 4, 8, 5, 9, 3);
  if (rc) { printf("theano run failed at step %d\n", rc); return rc; }
  rc = run(IAF_VARIANT_TF, 32, 64, 16, 16, 1);  // two row bands in the forward kernel, 64 x 64 weight-gradient tiles
  if (rc) { printf("tf run failed at step %d\n", rc); return 10 + rc; }
  printf("race_check ok\n");
  return 0;
}
