// tile_iter_check.cu -- host-side check of the tile enumeration shared by the roles of the symmetric kernels (csrc/sym_tc_dev.cuh
// Tile5Iter): for every row block I, every column split and a set of problem sizes,
//   * the run-based iterator (first_run / advance) visits exactly the tiles next_live visits, in the same order;
//   * diag(t) is true exactly when the tile's column block is the row block itself;
//   * over all row blocks and splits every unordered pair of 128-row blocks is visited once (the diagonal pair once, with diag set).
// Prints "ok <pairs checked>" or the first mismatch; exit status 0 / 1.  Built and run by tests/test_tile_iter_cpu.py.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sym_tc_dev.cuh"

using rpgp::tcdev::LiveRun;
using rpgp::tcdev::Tile5Iter;
using rpgp::tcdev::T5_ROWS;

int main() {
    long long checked = 0;
    const long long ns[] = {1, 100, 128, 129, 1000, 1025, 2500, 4000 + 55, 128 * 9, 128 * 10, 128 * 10 + 64, 128 * 33 + 31};
    for (long long n : ns) {
        const int B = (int)((n + T5_ROWS - 1) / T5_ROWS), half = B / 2 + 1;
        for (int nsplits : {1, 2, 3, 5, half}) {
            if (nsplits > half) continue;
            std::vector<int> seen((size_t)B * B, 0);
            const int per = (half + nsplits - 1) / nsplits;
            for (int I = 0; I < B; ++I)
                for (int by = 0; by < nsplits; ++by) {
                    Tile5Iter it;
                    it.I = I; it.B = B; it.n = n; it.k_begin = by * per;
                    it.ntiles = 4 * ((half < it.k_begin + per ? half : it.k_begin + per) - it.k_begin);
                    if (it.ntiles < 0) it.ntiles = 0;
                    LiveRun run = it.first_run();
                    for (int t = it.next_live(0); t < it.ntiles; t = it.next_live(t + 1), it.advance(run)) {
                        if (run.t != t) { printf("n=%lld B=%d I=%d split %d/%d: run-based iterator at %d, next_live at %d\n", n, B, I, by, nsplits, run.t, t); return 1; }
                        const int cb = it.block_of(it.k_begin + (t >> 2));
                        if (it.diag(t) != (cb == I)) { printf("n=%lld B=%d I=%d t=%d: diag %d but column block %d\n", n, B, I, t, (int)it.diag(t), cb); return 1; }
                        if ((t & 3) == 0) ++seen[(size_t)I * B + cb];
                        ++checked;
                    }
                    if (run.t != it.ntiles && run.t < it.ntiles) { printf("n=%lld B=%d I=%d: run-based iterator has tiles left (%d of %d)\n", n, B, I, run.t, it.ntiles); return 1; }
                }
            for (int a = 0; a < B; ++a)
                for (int b = a; b < B; ++b) {
                    const int cnt = seen[(size_t)a * B + b] + (a != b ? seen[(size_t)b * B + a] : 0);
                    if (cnt != 1) { printf("n=%lld B=%d nsplits=%d: block pair (%d, %d) visited %d times\n", n, B, nsplits, a, b, cnt); return 1; }
                }
        }
    }
    // the benchmark size (n = 1M: 7813 row blocks, the last one half full), the split counts the launchers choose there, sampled row blocks
    {
        const long long n = 1000000;
        const int B = (int)((n + T5_ROWS - 1) / T5_ROWS), half = B / 2 + 1;
        for (int nsplits : {1, 2, 3, 5})
            for (int I : {0, 1, 1953, 1954, 3905, 3906, 3907, 7811, 7812}) {
                const int per = (half + nsplits - 1) / nsplits;
                long long live_total = 0;
                for (int by = 0; by < nsplits; ++by) {
                    Tile5Iter it;
                    it.I = I; it.B = B; it.n = n; it.k_begin = by * per;
                    it.ntiles = 4 * ((half < it.k_begin + per ? half : it.k_begin + per) - it.k_begin);
                    if (it.ntiles < 0) it.ntiles = 0;
                    LiveRun run = it.first_run();
                    for (int t = it.next_live(0); t < it.ntiles; t = it.next_live(t + 1), it.advance(run)) {
                        if (run.t != t) { printf("n=1M I=%d split %d/%d: run-based iterator at %d, next_live at %d\n", I, by, nsplits, run.t, t); return 1; }
                        if (it.diag(t) != (it.block_of(it.k_begin + (t >> 2)) == I)) { printf("n=1M I=%d t=%d: diag mismatch\n", I, t); return 1; }
                        ++live_total;
                        ++checked;
                    }
                    if (run.t < it.ntiles) { printf("n=1M I=%d: run-based iterator has tiles left\n", I); return 1; }
                }
                // half column blocks of four tiles, minus the two dead tiles of the half-full last block when this row block meets it
                const bool meets_last = ((B - 1 - I + B) % B) < half;
                if (live_total != 4LL * half - (meets_last ? 2 : 0)) { printf("n=1M I=%d nsplits=%d: %lld live tiles\n", I, nsplits, live_total); return 1; }
            }
    }
    printf("ok %lld\n", checked);
    return 0;
}
