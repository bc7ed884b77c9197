// C++ test of the host-side mirror (include/zpcb200/zensim_b200.hpp), written the way the reference's own tests are
// (test/cuda/main.cu + test/utils/parallel_primitives.hpp:9-32: plain main(), throw on failure): reduce with
// getmax / getmin / plus over the reference's size list, then scan / radix_sort_pair against <algorithm>, then one
// composed APIC substep against the C oracle (test infrastructure, linked only here).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "zpcb200/zensim_b200.hpp"
extern "C" {
#include "../../oracle/oracle.h"
}
using namespace zsb200;

#define CHECK(cond, msg) do { if (!(cond)) throw std::runtime_error(std::string("FAILED: ") + msg); } while (0)

template <typename Op, typename HostOp> void test_reduction(CudaExecutionPolicy &pol, size_t n, int init, Op op, HostOp hop) {
  std::mt19937 rng((unsigned)n);
  std::vector<int> h(n);
  for (auto &v : h) v = (int)(rng() % 2000001) - 1000000;
  Vector<int> d(h), res(1);
  reduce(pol, d.begin(), d.end(), res.data(), init, op);
  int ref = init;
  for (int v : h) ref = hop(ref, v);
  CHECK(res.getVal() == ref, "reduce n=" + std::to_string(n));
}

int main() {
  auto pol = cuda_exec().device(0).sync(true);
  for (size_t n : {1ul, 2ul, 7ul, 16ul, 128ul, 1024ul, 2000000ul}) {  // test/cuda/main.cu:13-27
    test_reduction(pol, n, std::numeric_limits<int>::lowest(), getmax<int>{}, [](int a, int b) { return a > b ? a : b; });
    test_reduction(pol, n, std::numeric_limits<int>::max(), getmin<int>{}, [](int a, int b) { return a < b ? a : b; });
    test_reduction(pol, n, 0, plus<int>{}, [](int a, int b) { return a + b; });
  }
  {  // scans + stable radix sort, index exact
    const size_t n = 300007;
    std::mt19937 rng(7);
    std::vector<int> h(n);
    for (auto &v : h) v = (int)(rng() % 100);
    Vector<int> d(h), o(n);
    exclusive_scan(pol, d.begin(), d.end(), o.data());
    std::vector<int> ex(n);
    std::exclusive_scan(h.begin(), h.end(), ex.begin(), 0);
    CHECK(o.toHost() == ex, "exclusive_scan");
    inclusive_scan(pol, d.begin(), d.end(), o.data());
    std::inclusive_scan(h.begin(), h.end(), ex.begin());
    CHECK(o.toHost() == ex, "inclusive_scan");
    std::vector<uint32_t> k(n);
    std::vector<int> v(n);
    for (size_t i = 0; i < n; ++i) { k[i] = rng(); v[i] = (int)i; }
    Vector<uint32_t> dk(k), dko(n);
    Vector<int> dv(v), dvo(n);
    radix_sort_pair(pol, dk.data(), dv.data(), dko.data(), dvo.data(), n, 6, 24);  // the binning window
    std::vector<int> idx(v);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return ((k[a] >> 6) & 0x3ffff) < ((k[b] >> 6) & 0x3ffff); });
    CHECK(dvo.toHost() == idx, "radix_sort_pair values");
    auto ko = dko.toHost();
    for (size_t i = 0; i < n; ++i) CHECK(ko[i] == k[idx[i]], "radix_sort_pair keys");
    CHECK(pol.lastError() == 0, "latched error");
  }
  {  // composed substep vs the oracle
    const int s = 6, G = 32;
    const float dx = 1.f / G, dt = 1e-4f, gravity = -9.8f, vol = dx * dx * dx / 8.f;
    const size_t n = (size_t)8 * s * s * s;
    std::mt19937 rng(3);
    std::uniform_real_distribution<float> U(0.f, 1.f);
    std::vector<float> x(3 * n), v(3 * n), m(n), C(9 * n), F(9 * n);
    size_t p = 0;
    for (int i = 0; i < s; ++i) for (int j = 0; j < s; ++j) for (int k = 0; k < s; ++k) for (int q = 0; q < 8; ++q, ++p) {
      const int c[3] = {i, j, k};
      for (int d = 0; d < 3; ++d) {
        x[3 * p + d] = (7 + c[d] + (((q >> (2 - d)) & 1) + U(rng)) * 0.5f) * dx;
        v[3 * p + d] = (d == 1 ? -1.f : 0.f) + 0.2f * (U(rng) - 0.5f);
      }
      m[p] = 1000.f * vol;
      for (int d = 0; d < 9; ++d) { C[d + 9 * p] = 0.3f * (U(rng) - 0.5f); F[d + 9 * p] = (d % 4 == 0 ? 1.f : 0.f) + 0.04f * (U(rng) - 0.5f); }
    }
    Particles pars(n);
    cudaMemcpy(pars.X.data(), x.data(), 12 * n, cudaMemcpyHostToDevice);
    cudaMemcpy(pars.V.data(), v.data(), 12 * n, cudaMemcpyHostToDevice);
    cudaMemcpy(pars.M.data(), m.data(), 4 * n, cudaMemcpyHostToDevice);
    cudaMemcpy(pars.C.data(), C.data(), 36 * n, cudaMemcpyHostToDevice);
    cudaMemcpy(pars.F.data(), F.data(), 36 * n, cudaMemcpyHostToDevice);
    HashTable table(n / 8);
    FixedCorotatedConfig model;
    model.volume = vol;
    pol(PartitionForParticles{pars, dx, table});
    const int nb = table.size();
    Grids grids(dx, nb);
    Vector<float> maxVel(1);
    maxVel.setVal(0.f);
    pol(CleanGridBlocks{grids, table});
    pol(P2GTransfer{dt, model, pars, table, grids});
    auto g1 = grids.blocks.toHost();
    pol(ComputeGridBlockVelocity{grids, table, dt, gravity, maxVel.data(), 1});
    pol(G2PTransfer{dt, grids, table, pars});
    CHECK(pol.lastError() == 0, "latched error in substep");
    // oracle on the table the GPU built
    auto hk = table.keys.toHost(), hi = table.indices.toHost();
    std::vector<float> og((size_t)nb * 448, 0.f);
    zo_p2g_fcr((int)n, x.data(), v.data(), m.data(), C.data(), F.data(), dx, dt, model.E, model.nu, vol, table._tableSize, hk.data(), hi.data(), og.data());
    double mass = 0, scale[7] = {0}, err[7] = {0};
    for (size_t i = 0; i < og.size(); ++i) {
      const int ch = (int)((i / 64) % 7);
      scale[ch] = std::max(scale[ch], (double)std::fabs(og[i]));
      err[ch] = std::max(err[ch], (double)std::fabs(og[i] - g1[i]));
      if (ch == 0) mass += g1[i];
    }
    for (int ch = 0; ch < 7; ++ch) CHECK(err[ch] <= (ch < 4 ? 1e-5 : 1e-4) * scale[ch], "p2g channel " + std::to_string(ch));
    CHECK(std::fabs(mass / (n * 1000.0 * vol) - 1) < 1e-5, "mass conservation");
    float mx = 0.f;
    const float extf[3] = {0.f, gravity, 0.f};
    zo_grid_update(nb, og.data(), dt, extf, 1, &mx);
    zo_g2p((int)n, x.data(), v.data(), C.data(), F.data(), dx, dt, table._tableSize, hk.data(), hi.data(), og.data());
    CHECK(std::fabs(maxVel.getVal() - mx) <= 1e-5f * mx, "maxVel");
    auto gx = pars.X.toHost(), gv = pars.V.toHost(), gF = pars.F.toHost();
    double ex = 0, ev = 0, eF = 0;
    for (size_t i = 0; i < 3 * n; ++i) { ex = std::max(ex, (double)std::fabs(gx[i] - x[i])); ev = std::max(ev, (double)std::fabs(gv[i] - v[i])); }
    for (size_t i = 0; i < 9 * n; ++i) eF = std::max(eF, (double)std::fabs(gF[i] - F[i]));
    CHECK(ex <= 1e-5 && ev <= 1.2e-5 && eF <= 1.1e-5, "g2p x/v/F");
    std::printf("substep ok: %d blocks, err x %.2e v %.2e F %.2e\n", nb, ex, ev, eF);

    // the same substep on a SparseGrid<3,f32,8> (bht partition, side-8 blocks): same particles within rounding
    Particles pars2(n);
    {
      // re-create the initial state (x, v, C, F were advanced by the oracle above: rebuild them with the same generator)
      std::mt19937 rng2(3);
      std::vector<float> x2(3 * n), v2(3 * n), C2(9 * n), F2(9 * n);
      size_t q2 = 0;
      for (int i = 0; i < s; ++i) for (int j = 0; j < s; ++j) for (int k = 0; k < s; ++k) for (int q = 0; q < 8; ++q, ++q2) {
        const int c3[3] = {i, j, k};
        for (int d = 0; d < 3; ++d) {
          x2[3 * q2 + d] = (7 + c3[d] + (((q >> (2 - d)) & 1) + U(rng2)) * 0.5f) * dx;
          v2[3 * q2 + d] = (d == 1 ? -1.f : 0.f) + 0.2f * (U(rng2) - 0.5f);
        }
        for (int d = 0; d < 9; ++d) { C2[d + 9 * q2] = 0.3f * (U(rng2) - 0.5f); F2[d + 9 * q2] = (d % 4 == 0 ? 1.f : 0.f) + 0.04f * (U(rng2) - 0.5f); }
      }
      cudaMemcpy(pars2.X.data(), x2.data(), 12 * n, cudaMemcpyHostToDevice);
      cudaMemcpy(pars2.V.data(), v2.data(), 12 * n, cudaMemcpyHostToDevice);
      cudaMemcpy(pars2.M.data(), m.data(), 4 * n, cudaMemcpyHostToDevice);
      cudaMemcpy(pars2.C.data(), C2.data(), 36 * n, cudaMemcpyHostToDevice);
      cudaMemcpy(pars2.F.data(), F2.data(), 36 * n, cudaMemcpyHostToDevice);
    }
    SparseGrid sg(7, n / 16);
    sg.scale(dx);
    pol(SgPartitionForParticles{pars2, sg});
    CHECK(sg._table._buildSuccess.getVal() == 1 && sg.numBlocks() > 0, "bht build");
    maxVel.setVal(0.f);
    pol(SgCleanGridBlocks{sg});
    pol(SgP2GTransfer{dt, model, pars2, sg});
    pol(SgComputeGridBlockVelocity{sg, dt, gravity, maxVel.data(), 1});
    pol(SgG2PTransfer{dt, sg, pars2});
    CHECK(pol.lastError() == 0, "latched error in SparseGrid substep");
    CHECK(std::fabs(maxVel.getVal() - mx) <= 1e-5f * mx, "SparseGrid maxVel");
    auto sx = pars2.X.toHost(), sv = pars2.V.toHost(), sF = pars2.F.toHost();
    double dxm = 0, dvm = 0, dFm = 0;
    for (size_t i = 0; i < 3 * n; ++i) { dxm = std::max(dxm, (double)std::fabs(sx[i] - x[i])); dvm = std::max(dvm, (double)std::fabs(sv[i] - v[i])); }
    for (size_t i = 0; i < 9 * n; ++i) dFm = std::max(dFm, (double)std::fabs(sF[i] - F[i]));
    CHECK(dxm <= 1e-5 && dvm <= 1.2e-5 && dFm <= 1.1e-5, "SparseGrid g2p x/v/F");
    std::printf("SparseGrid substep ok: %zu blocks of 8^3, err x %.2e v %.2e F %.2e\n", sg.numBlocks(), dxm, dvm, dFm);

    // the block-binned fast path through the mirror, on both grids: bin -> P2G -> update -> G2P -> unbin; binned particle i is
    // AoS particle order[i], so the oracle's result (x, v, F above) is compared through that permutation
    for (int which = 0; which < 2; ++which) {
      Particles pars3(n), back(n);
      {
        std::mt19937 rng3(3);
        std::vector<float> x3(3 * n), v3(3 * n), C3(9 * n), F3(9 * n);
        size_t q3 = 0;
        for (int i = 0; i < s; ++i) for (int j = 0; j < s; ++j) for (int k = 0; k < s; ++k) for (int q = 0; q < 8; ++q, ++q3) {
          const int c3[3] = {i, j, k};
          for (int d = 0; d < 3; ++d) {
            x3[3 * q3 + d] = (7 + c3[d] + (((q >> (2 - d)) & 1) + U(rng3)) * 0.5f) * dx;
            v3[3 * q3 + d] = (d == 1 ? -1.f : 0.f) + 0.2f * (U(rng3) - 0.5f);
          }
          for (int d = 0; d < 9; ++d) { C3[d + 9 * q3] = 0.3f * (U(rng3) - 0.5f); F3[d + 9 * q3] = (d % 4 == 0 ? 1.f : 0.f) + 0.04f * (U(rng3) - 0.5f); }
        }
        cudaMemcpy(pars3.X.data(), x3.data(), 12 * n, cudaMemcpyHostToDevice);
        cudaMemcpy(pars3.V.data(), v3.data(), 12 * n, cudaMemcpyHostToDevice);
        cudaMemcpy(pars3.M.data(), m.data(), 4 * n, cudaMemcpyHostToDevice);
        cudaMemcpy(pars3.C.data(), C3.data(), 36 * n, cudaMemcpyHostToDevice);
        cudaMemcpy(pars3.F.data(), F3.data(), 36 * n, cudaMemcpyHostToDevice);
      }
      Vector<int> order(n);
      maxVel.setVal(0.f);
      int nbins = 0, status = 0;
      if (which == 0) {
        HashTable table3(n / 8);
        pol(PartitionForParticles{pars3, dx, table3});
        Grids grids3(dx, table3.size());
        BinnedParticles bins((size_t)n, 2 * table3.size() + 64);
        pol(BinParticles{pars3, table3, dx, bins, order.data()});
        pol(CleanGridBlocks{grids3, table3});
        pol(P2GTransferBinned{dt, model, bins, table3, grids3});
        pol(ComputeGridBlockVelocity{grids3, table3, dt, gravity, maxVel.data(), 1});
        pol(G2PTransferBinned{dt, grids3, table3, bins});
        pol(UnbinParticles{bins, back});
        nbins = bins.bins(); status = bins.statusWord();
      } else {
        SparseGrid sg3(7, n / 16);
        sg3.scale(dx);
        pol(SgPartitionForParticles{pars3, sg3});
        BinnedParticles bins((size_t)n, 8 * (int)sg3.numBlocks() + 64);
        pol(SgBinParticles{pars3, sg3, bins, order.data()});
        pol(SgCleanGridBlocks{sg3});
        pol(SgP2GTransferBinned{dt, model, bins, sg3});
        pol(SgComputeGridBlockVelocity{sg3, dt, gravity, maxVel.data(), 1});
        pol(SgG2PTransferBinned{dt, sg3, bins});
        pol(UnbinParticles{bins, back});
        nbins = bins.bins(); status = bins.statusWord();
      }
      CHECK(pol.lastError() == 0 && status == 0 && nbins > 0, "latched error / status word on the binned path");
      CHECK(std::fabs(maxVel.getVal() - mx) <= 1e-5f * mx, "binned maxVel");
      auto ord = order.toHost();
      auto bx = back.X.toHost(), bv = back.V.toHost(), bF = back.F.toHost();
      double e1 = 0, e2 = 0, e3 = 0;
      for (size_t i = 0; i < n; ++i) {
        const size_t j = (size_t)ord[i];
        for (int d = 0; d < 3; ++d) { e1 = std::max(e1, (double)std::fabs(bx[3 * i + d] - x[3 * j + d])); e2 = std::max(e2, (double)std::fabs(bv[3 * i + d] - v[3 * j + d])); }
        for (int d = 0; d < 9; ++d) e3 = std::max(e3, (double)std::fabs(bF[9 * i + d] - F[9 * j + d]));
      }
      CHECK(e1 <= 1e-5 && e2 <= 1.2e-5 && e3 <= 1.1e-5, "binned g2p x/v/F");
      std::printf("binned fast path on %s ok: %d bins, err x %.2e v %.2e F %.2e\n", which ? "SparseGrid<3,f32,8>" : "Grids<f32,3,4>", nbins, e1, e2, e3);
    }
  }
  {  // merge_sort_pair: stable, in place, float keys with signed zeros
    const size_t n = 100003;
    std::mt19937 rng(5);
    std::vector<float> k(n);
    std::vector<int> v(n);
    for (size_t i = 0; i < n; ++i) { k[i] = (float)((int)(rng() % 41) - 20) * 0.5f; v[i] = (int)i; }
    k[3] = -0.0f; k[7] = 0.0f;
    Vector<float> dk(k);
    Vector<int> dv(v);
    merge_sort_pair(pol, dk.data(), dv.data(), n);
    std::vector<int> idx(v);
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return k[a] < k[b]; });
    CHECK(dv.toHost() == idx, "merge_sort_pair values");
    auto ko = dk.toHost();
    for (size_t i = 0; i < n; ++i) CHECK(std::memcmp(&ko[i], &k[idx[i]], 4) == 0, "merge_sort_pair keys");
  }
  std::printf("all host-mirror tests passed\n");
  return 0;
}
