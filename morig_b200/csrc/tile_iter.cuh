// Tile streams of the persistent tcgen05 kernels (gemm_tc.cuh): which (n-tile, row tile, key-frame) a CTA works on.
// Plain integer arithmetic, kept in its own header so that the host unit test (tests/host/tile_iter_test.cpp, built
// with g++ by tests/test_host_logic.py) exercises exactly the code the kernels run.
#pragma once

#ifdef __CUDACC__
#define MORIG_HD __host__ __device__ __forceinline__
#else
#define MORIG_HD inline
#endif

namespace morig {
namespace tc {

constexpr int BM = 128;                      // rows (edges / vertices) per CTA tile

struct TileCoord { int n_tile, m0, frame; };

// tile stream of one CTA: t = first, first + step, ... < total;  m0 = ((r % ntm) * mult + rank) * BM
struct TileMap {
    int ntn, ntm, total, first, step, mult, rank;
    MORIG_HD TileCoord decode(int t) const {
        TileCoord c;
        c.n_tile = t % ntn;
        const int r = t / ntn;
        c.m0 = ((r % ntm) * mult + rank) * BM;
        c.frame = r / ntm;
        return c;
    }
    MORIG_HD int my_tiles() const { return total > first ? (total - 1 - first) / step + 1 : 0; }
};

// The same stream with the coordinates kept incrementally: one add / compare / select chain per tile instead of the
// four integer divisions of decode() (ncu: ~190 of a producer warp's ~880 instructions per tile were tile arithmetic).
struct TileIter {
    int t, n_tile, mi, frame;        // mi = m-tile (cta_group::2: pair-tile) index inside the key-frame
    int d_n, d_m, d_f;               // `step` decomposed in the mixed radix (ntn, ntm)
    int ntn, ntm, total, step, mult, rank;
    MORIG_HD void init(const TileMap &tm) {
        ntn = tm.ntn; ntm = tm.ntm; total = tm.total; step = tm.step; mult = tm.mult; rank = tm.rank;
        t = tm.first;
        n_tile = t % ntn;
        const int r = t / ntn;
        mi = r % ntm; frame = r / ntm;
        d_n = step % ntn;
        const int a = step / ntn;
        d_m = a % ntm; d_f = a / ntm;
    }
    MORIG_HD bool valid() const { return t < total; }
    MORIG_HD int m0() const { return (mi * mult + rank) * BM; }
    MORIG_HD void next() {
        t += step;
        n_tile += d_n;
        int c = (n_tile >= ntn) ? 1 : 0;
        n_tile -= c ? ntn : 0;
        mi += d_m + c;
        c = (mi >= ntm) ? 1 : 0;
        mi -= c ? ntm : 0;
        frame += d_f + c;
    }
};

}  // namespace tc
}  // namespace morig
