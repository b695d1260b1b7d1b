// Shared-memory placement of the state vector for the recurrence kernels (host code).
//
// A recurrence step gathers, for every unit, the states its row of W references: `ld.shared.f64` with 32 unrelated
// addresses per warp.  With the natural placement (state j in 8-byte slot j) a half-warp of 16 such addresses hits
// some bank pair 3 or more times, so a gather costs ~3.3 wavefronts instead of 2 and the shared-memory pipe -- not
// arithmetic -- bounds the phase (16 warps x 20 gathers x 3.3 = ~900 of the ~1300 cycles measured at N=2000, 500 units
// per CTA, profiles/recur_trace_r02_after.txt).  But the placement is ours to choose: any permutation `slot[j]`
// works as long as writers and readers agree.  The conflicts of a kernel are known in advance -- lane l of warp w
// reads column indices[indptr[row] + k] in round k -- so this file picks the bank (slot mod 16) of every state by a
// local search over those access sets (swaps of two states inside a warp's run of 32 slots, see plan_bank_layout).
// Nothing about the arithmetic changes (same values, same order): the results are bit-identical, only the addresses
// move.  Measured: gather wavefronts per CTA and step 627 -> 456 at N=2000, the recurrence step 4 % shorter.
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "kernels.cuh"

namespace xesn {

namespace {
constexpr int NB = 16;  // 8-byte bank pairs of a half-warp access

struct AccessSets {
    // set s = the distinct columns one half-warp reads in one gather instruction (+ the zero slot if a lane is padded)
    std::vector<int> set_ptr, set_cols;      // CSR: columns of each set
    std::vector<unsigned char> has_zero;     // the set also reads the always-zero slot
    std::vector<int> col_ptr, col_sets;      // transpose: sets of each column
};

// rows [u0, u1) of every CTA are dealt to threads in order (unit = u0 + tid); a warp runs `kreg` rounds plus rounds of
// `ktail` while its longest row lasts -- the access pattern shared by recur_cluster.cuh, recur_body.cuh, predict_flow.cu
AccessSets build_sets(int N, const int* indptr, const int* indices, int upc, int kreg, int ktail) {
    AccessSets a;
    a.set_ptr.push_back(0);
    std::vector<int> tmp;
    for (int u0 = 0; u0 < N; u0 += upc) {
        const int u1 = std::min(N, u0 + upc);
        for (int w0 = u0; w0 < u1; w0 += 32) {
            const int w1 = std::min(u1, w0 + 32);
            int maxlen = 0;
            for (int r = w0; r < w1; ++r) maxlen = std::max(maxlen, indptr[r + 1] - indptr[r]);
            int rounds = kreg;
            if (maxlen > kreg) rounds += (maxlen - kreg + ktail - 1) / ktail * ktail;
            for (int h0 = w0; h0 < w1; h0 += 16) {
                const int h1 = std::min(w1, h0 + 16);
                for (int k = 0; k < rounds; ++k) {
                    tmp.clear();
                    bool zero = false;
                    for (int r = h0; r < h1; ++r) {
                        if (k < indptr[r + 1] - indptr[r]) tmp.push_back(indices[indptr[r] + k]);
                        else zero = true;
                    }
                    std::sort(tmp.begin(), tmp.end());
                    tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
                    if (tmp.empty()) continue;   // a broadcast of the zero slot: one wavefront whatever the placement
                    a.set_cols.insert(a.set_cols.end(), tmp.begin(), tmp.end());
                    a.set_ptr.push_back((int)a.set_cols.size());
                    a.has_zero.push_back(zero ? 1 : 0);
                }
            }
        }
    }
    const int ns = (int)a.has_zero.size();
    a.col_ptr.assign(N + 1, 0);
    for (int c : a.set_cols) ++a.col_ptr[c + 1];
    for (int j = 0; j < N; ++j) a.col_ptr[j + 1] += a.col_ptr[j];
    a.col_sets.resize(a.set_cols.size());
    std::vector<int> fill(a.col_ptr.begin(), a.col_ptr.end() - 1);
    for (int s = 0; s < ns; ++s)
        for (int e = a.set_ptr[s]; e < a.set_ptr[s + 1]; ++e) a.col_sets[fill[a.set_cols[e]]++] = s;
    return a;
}

inline int max16(const unsigned char* c) {
    int m = 0;
    for (int b = 0; b < NB; ++b) m = std::max(m, (int)c[b]);
    return m;
}
}  // namespace

// slot_out[j]: 8-byte slot of state j inside a copy of the state vector; the always-zero slot sits at index `zero_slot`.
// The placement is a permutation INSIDE every warp's own run of 32 natural slots (units u0 + 32 w ... of a CTA): the
// peers receive a warp's new states as remote stores, and those only stay cheap while they fall into one 256-byte
// segment (tools/micro/dsmem_allgather.cu: a placement scattered over the whole slice makes the exchange 2.4x slower,
// a permutation inside the segments costs 3 %).  Inside a run every bank pair occurs exactly twice, so the search
// swaps two states of a run whenever that lowers the wavefront count of the gathers they take part in.
// Returns the gather wavefronts per step summed over all CTAs after the search; *before = natural placement.
long long plan_bank_layout(int N, const int* indptr, const int* indices, int upc, int kreg, int ktail, int n_slots,
                           int zero_slot, int* slot_out, long long* before) {
    (void)n_slots;
    const AccessSets a = build_sets(N, indptr, indices, upc, kreg, ktail);
    const int ns = (int)a.has_zero.size();
    const int zero_bank = zero_slot % NB;
    std::vector<unsigned char> cnt((size_t)ns * NB, 0), smax(ns, 0);
    std::vector<int> slot(N), run_of(N), run_lo, run_hi;
    for (int u0 = 0; u0 < N; u0 += upc)
        for (int w0 = u0; w0 < std::min(N, u0 + upc); w0 += 32) {
            const int w1 = std::min(std::min(N, u0 + upc), w0 + 32);
            for (int j = w0; j < w1; ++j) run_of[j] = (int)run_lo.size();
            run_lo.push_back(w0), run_hi.push_back(w1);
        }
    for (int j = 0; j < N; ++j) slot[j] = j;
    long long total = 0;
    for (int s = 0; s < ns; ++s) {
        unsigned char* c = &cnt[(size_t)s * NB];
        if (a.has_zero[s]) ++c[zero_bank];
        for (int e = a.set_ptr[s]; e < a.set_ptr[s + 1]; ++e) ++c[slot[a.set_cols[e]] % NB];
        smax[s] = (unsigned char)max16(c);
        total += smax[s];
    }
    if (before) *before = total;
    auto sets_cost = [&](int j) {
        int c = 0;
        for (int e = a.col_ptr[j]; e < a.col_ptr[j + 1]; ++e) c += smax[a.col_sets[e]];
        return c;
    };
    auto move = [&](int j, int from, int to) {   // state j changes its bank; keeps cnt and smax current
        for (int e = a.col_ptr[j]; e < a.col_ptr[j + 1]; ++e) {
            const int s = a.col_sets[e];
            unsigned char* c = &cnt[(size_t)s * NB];
            --c[from], ++c[to];
            total -= smax[s];
            smax[s] = (unsigned char)max16(c);
            total += smax[s];
        }
    };
    std::vector<int> order(N);
    std::iota(order.begin(), order.end(), 0);
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    auto next = [&rng]() { rng ^= rng << 13, rng ^= rng >> 7, rng ^= rng << 17; return rng; };
    for (int pass = 0; pass < 12; ++pass) {
        for (int i = N - 1; i > 0; --i) std::swap(order[i], order[next() % (uint64_t)(i + 1)]);
        const long long at_start = total;
        for (int j : order) {
            if (a.col_ptr[j] == a.col_ptr[j + 1] || sets_cost(j) == a.col_ptr[j + 1] - a.col_ptr[j]) continue;  // conflict-free
            const int r = run_of[j];
            int best_k = -1;
            long long best_total = total;
            for (int k = run_lo[r]; k < run_hi[r]; ++k) {
                const int bj = slot[j] % NB, bk = slot[k] % NB;
                if (k == j || bj == bk) continue;
                move(j, bj, bk), move(k, bk, bj);
                if (total < best_total) best_total = total, best_k = k;
                move(k, bj, bk), move(j, bk, bj);   // undo
            }
            if (best_k >= 0) {
                const int bj = slot[j] % NB, bk = slot[best_k] % NB;
                move(j, bj, bk), move(best_k, bk, bj);
                std::swap(slot[j], slot[best_k]);
            }
        }
        if (total == at_start) break;
    }
    for (int j = 0; j < N; ++j) slot_out[j] = slot[j];
    return total;
}

}  // namespace xesn
