// Host unit test of the tile-stream arithmetic the tcgen05 kernels use (morig_b200/csrc/tile_iter.cuh): the incremental
// TileIter must visit exactly the coordinates TileMap::decode gives for t = first, first + step, ... and my_tiles() must
// count them.  Built and run by tests/test_host_logic.py (g++, no CUDA needed).
#include <cstdio>
#include <cstdlib>
#include "../../morig_b200/csrc/tile_iter.cuh"

int main() {
    using namespace morig::tc;
    unsigned seed = 12345u;
    auto rnd = [&](int lo, int hi) { seed = seed * 1664525u + 1013904223u; return lo + (int)((seed >> 8) % (unsigned)(hi - lo + 1)); };
    long long checked = 0;
    for (int trial = 0; trial < 20000; ++trial) {
        TileMap tm;
        tm.ntn = rnd(1, 9); tm.ntm = rnd(1, 700); const int frames = rnd(1, 6);
        tm.total = tm.ntn * tm.ntm * frames;
        tm.step = rnd(1, 160); tm.first = rnd(0, tm.step - 1);
        tm.mult = rnd(1, 2); tm.rank = tm.mult == 2 ? rnd(0, 1) : 0;
        TileIter it;
        it.init(tm);
        int count = 0;
        for (int t = tm.first; t < tm.total; t += tm.step, ++count) {
            const TileCoord c = tm.decode(t);
            if (!it.valid() || it.t != t || it.n_tile != c.n_tile || it.m0() != c.m0 || it.frame != c.frame) {
                std::printf("MISMATCH trial %d t %d: iter (%d %d %d %d) decode (%d %d %d)\n", trial, t, it.t, it.n_tile, it.m0(), it.frame,
                            c.n_tile, c.m0, c.frame);
                return 1;
            }
            it.next();
            ++checked;
        }
        if (it.valid() || count != tm.my_tiles()) {
            std::printf("END MISMATCH trial %d: valid %d count %d my_tiles %d\n", trial, (int)it.valid(), count, tm.my_tiles());
            return 1;
        }
    }
    std::printf("OK %lld tiles\n", checked);
    return 0;
}
