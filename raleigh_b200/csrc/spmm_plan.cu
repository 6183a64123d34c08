// Host-side set-up for the block SpMM (no device work): order the 32-row runs of a CSR
// matrix so that the runs one CTA processes together gather from overlapping column
// segments.
//
// Why: with vector-major block vectors the SpMM kernel maps lane = row, so a warp (a run of
// 32 consecutive rows) reads, for every vector, one 256-byte segment of X per stencil
// neighbour.  On a 3-D grid numbered lexicographically the neighbours at +-N and +-N^2 belong
// to other runs; with CTAs made of CONSECUTIVE runs those lines are fetched from L2 again by
// whichever SM owns the neighbouring run (measured on B200: ~5x the X block through L2 for
// the 7-point Laplacian, which is what bounds the kernel at ~55 % of HBM peak).  Grouping
// runs that are neighbours in y and z into one CTA turns those re-reads into L1 hits:
// 2x2 runs -> 3x, 4x2 -> 2.5x, 4x4 -> 2x.
//
// The grouping is generic (no geometry is assumed): the run graph of a structurally
// symmetric matrix is walked greedily, a cluster growing by the unassigned run that shares
// the most column segments with what the cluster already gathers.
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace rl {

namespace {

// F[a]: distinct 32-column segments (restricted to [0, nruns)) that run a references with
// at least `thresh` entries, as a CSR-like pair (fptr, fidx).
void run_footprints(int64_t nrows, const int64_t* indptr, const int32_t* indices, int thresh,
                    std::vector<int64_t>& fptr, std::vector<int32_t>& fidx) {
    const int64_t nruns = (nrows + 31) / 32;
    fptr.assign(nruns + 1, 0);
    fidx.clear();
    std::vector<int32_t> segs;
    for (int64_t a = 0; a < nruns; ++a) {
        const int64_t r0 = a * 32, r1 = std::min<int64_t>(r0 + 32, nrows);
        segs.clear();
        for (int64_t p = indptr[r0]; p < indptr[r1]; ++p) {
            const int64_t s = indices[p] >> 5;
            if (s < nruns) segs.push_back((int32_t)s);
        }
        std::sort(segs.begin(), segs.end());
        for (size_t i = 0; i < segs.size();) {
            size_t j = i;
            while (j < segs.size() && segs[j] == segs[i]) ++j;
            if ((int)(j - i) >= thresh) fidx.push_back(segs[i]);
            i = j;
        }
        fptr[a + 1] = (int64_t)fidx.size();
    }
}

}  // namespace

}  // namespace rl

using namespace rl;

extern "C" {

int rl_spmm_cluster_runs(int64_t nrows, const int64_t* indptr_h, const int32_t* indices_h, int group,
                         int32_t* order_out_h, double* footprint_ratio_out) {
    if (nrows < 0 || group < 1 || (nrows > 0 && (!indptr_h || !indices_h || !order_out_h))) return RL_E_ARG;
    const int64_t nruns = (nrows + 31) / 32;
    if (nruns > INT32_MAX) return RL_E_ARG;
    std::vector<int64_t> fptr;
    std::vector<int32_t> fidx;
    run_footprints(nrows, indptr_h, indices_h, 8, fptr, fidx);

    std::vector<char> assigned(nruns, 0);
    std::vector<int32_t> stamp(nruns, -1);       // stamp[s] == cid: segment s is in the current cluster's footprint
    std::vector<int32_t> cand;                   // unassigned runs adjacent to the cluster
    std::vector<int32_t> cand_stamp(nruns, -1);
    int64_t out = 0, next_free = 0, footprint_total = 0;
    int32_t cid = 0;
    for (int64_t seed = 0; seed < nruns; ++seed) {
        if (assigned[seed]) continue;
        cand.clear();
        int members = 0;
        int64_t cluster_fp = 0;
        auto add = [&](int32_t a) {
            assigned[a] = 1;
            order_out_h[out++] = a;
            ++members;
            for (int64_t q = fptr[a]; q < fptr[a + 1]; ++q) {
                const int32_t s = fidx[q];
                if (stamp[s] != cid) { stamp[s] = cid; ++cluster_fp; }
                if (!assigned[s] && cand_stamp[s] != cid) { cand_stamp[s] = cid; cand.push_back(s); }
            }
        };
        add((int32_t)seed);
        while (members < group) {
            int best = -1, best_score = 0;
            for (size_t i = 0; i < cand.size(); ++i) {
                const int32_t c = cand[i];
                if (assigned[c]) continue;
                int score = 0;
                for (int64_t q = fptr[c]; q < fptr[c + 1]; ++q) score += stamp[fidx[q]] == cid;
                if (score > best_score || (score == best_score && best >= 0 && c < best)) { best = c; best_score = score; }
            }
            if (best < 0) {
                // nothing adjacent is left: fill the CTA with the lowest free run so every CTA stays full
                while (next_free < nruns && assigned[next_free]) ++next_free;
                if (next_free >= nruns) break;
                best = (int32_t)next_free;
            }
            add(best);
        }
        footprint_total += cluster_fp;
        ++cid;
    }
    if (footprint_ratio_out) *footprint_ratio_out = nruns > 0 ? (double)footprint_total / (double)nruns : 0.0;
    return out == nruns ? 0 : RL_E_ARG;
}

}  // extern "C"
