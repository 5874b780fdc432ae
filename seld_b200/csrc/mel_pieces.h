// Host-side construction of the "piece" form of a triangular mel bank (used by seld_plan_create and by the CPU
// emulation in tests/emu).  Input: dense [n_bins][n_mels] float32 table with at most two non-zeros per row, in adjacent
// filters.  Lane u of a frame team (kTeamLanes = 64 threads: the two warps that share one STFT frame) owns bins
// [u*bpt, (u+1)*bpt); a piece is a maximal run of one lane's bins feeding the same filter pair (seg, seg + 1).  See
// Tables in extract_core.cuh for how the kernels consume these arrays.
#pragma once

#include <string>
#include <vector>

namespace seld {

constexpr int kTeamLanes = 64;

struct MelPieces {
    int bpt = 0;                                  // bins per lane
    std::vector<float> w01;                       // [64*bpt][2]  0.25 * (w into seg, w into seg + 1)
    std::vector<unsigned long long> endmask;      // [64]
    std::vector<int> slot0;                       // [64] record slot of the lane's first piece
    std::vector<int> slot1;                       // [64] record slot of its second piece; later pieces follow at +1
    std::vector<int> pb;                          // [n_mels + 2]  (compact layout only)
    int n_pieces = 0;
    int max_pieces_per_seg = 0;
    // Record layout.  compact: slot = piece index (pieces sorted by segment; the gather walks pb[]).
    // seg_major (n_mels <= 64, <= kSegMajorRanks pieces per segment, no empty segment inside a lane's run): the first two
    // pieces of segment s sit at slots s and s + kSegMajorPitch, so gather lane u reads slots u and u + 65 -- consecutive
    // lanes hit consecutive records (bank-conflict free).  The few segments with a third / fourth piece (the widest
    // filters) put them in an overflow area behind slot 130 and name them in ov[s] (low / high 16 bits); every other
    // entry of ov points at slot kSegMajorZero, which is never written and therefore stays zero.
    bool seg_major = false;
    int n_slots = 0;
    std::vector<int> ov;                          // [64]
    // ---- flush-free "lane" form (FOA, n_fft 1024): lane u owns bins [lane_beg[u], lane_beg[u] + bpt) -- at most bpt of them count,
    // cut so that a lane touches at most kLaneSegs consecutive segments, hence at most kLaneSegs + 1 = 4 consecutive filters
    // seg0 .. seg0 + 3.  It accumulates those four partial sums per channel as two packed pairs with per-bin weights
    // w4 = (a0, a1, b0, b1), no flush inside the bin loop, and stores ONE record; filter m then adds the few records gtab[m] names.
    bool lanes_ok = false;
    std::vector<int> lane_beg;                    // [64]
    std::vector<int> lane_seg0;                   // [64]
    std::vector<float> w4;                        // [64 * bpt][4]  0.25 * weights into filters seg0, seg0 + 1 | seg0 + 2, seg0 + 3
    std::vector<int> gtab;                        // [64][kLaneGatherMax] word offset (lane * kLaneRecWords + j) of the t-th record of filter m
    int gather_n[2] = {0, 0};                     // entries the filters of team warp 0 / 1 need at most
    int lane_gather_wavefronts = 0;               // modelled shared-memory wavefronts of one channel of the gather (ideal: gather_n[0] + gather_n[1])
    bool lane_spread = false;                     // every 16-lane group starts its runs at 16 distinct residues mod 16
};

constexpr int kLaneSegs = 3;
constexpr int kLaneRecWords = 30;                 // 7 channels x (F0, F1, F2, F3) = 28 words, padded to an even, bank-friendly pitch
constexpr int kLaneGatherMax = 8;
constexpr int kLaneZeroRec = 64;                  // record 64 is never written: absent table entries point at it

constexpr int kSegMajorRanks = 4;
constexpr int kSegMajorPitch = 65;      // odd, so the pieces of one segment land in different bank groups
constexpr int kSegMajorZero = 64 + kSegMajorPitch;      // slot 129 = rank 1 of the non-existent segment 64

// ---------------------------------------------------------------- lane form of the bank (FOA, n_fft 1024)
// Shared-memory model the layout is tuned for (checked against ncu's per-instruction wavefront counts on B200): a 64-bit
// access is served 16 lanes at a time and is conflict-free iff the 16 word pairs fall into distinct bank pairs; a 32-bit
// access is served 32 lanes at a time, one wavefront per distinct address that shares a bank.  Correctness never depends on
// it -- only the number of wavefronts does.
struct LaneRun { int beg, n, s0; };

// cut the bin axis into runs of <= bpt bins over <= kLaneSegs consecutive segments; with `spread`, no more than four runs may
// start at the same bin index mod 16 (so that every group of 16 lanes can be given 16 distinct residues: conflict-free
// 64-bit reads of S[beg + i] and S[N - beg - i])
inline bool cut_lane_runs(const std::vector<int>& seg, int n_bins, int bpt, bool spread, std::vector<LaneRun>& runs) {
    constexpr int TL = kTeamLanes;
    auto longest = [&](int k, int& s0) {
        int n = 0; s0 = -1;
        while (k + n < n_bins && n < bpt) {
            const int sg = seg[k + n];
            if (sg >= 0) { if (s0 < 0) s0 = sg; if (sg - s0 >= kLaneSegs) break; }
            ++n;
        }
        return n;
    };
    int used[16] = {0};
    long budget = 200000;
    runs.clear();
    // depth-first, longest run first
    struct Frame { int k, n; };
    std::vector<Frame> st;
    int k = 0;
    for (;;) {
        if (k >= n_bins) return true;
        bool placed = false;
        if (int(st.size()) < TL && (!spread || used[k & 15] < 4) && budget-- > 0) {
            int s0; const int n = longest(k, s0);
            if (n > 0 && n_bins - (k + n) <= (TL - int(st.size()) - 1) * bpt) {
                ++used[k & 15]; st.push_back({k, n}); runs.push_back({k, n, s0 < 0 ? 0 : s0}); k += n; placed = true;
            }
        }
        if (placed) continue;
        // back-track: shorten the most recent run that can still be shortened
        for (;;) {
            if (st.empty() || budget <= 0) return false;
            Frame& f = st.back();
            if (spread && f.n > 1 && n_bins - (f.k + f.n - 1) <= (TL - int(st.size())) * bpt) {
                --f.n; runs.back().n = f.n; k = f.k + f.n; break;
            }
            --used[f.k & 15]; st.pop_back(); runs.pop_back();
        }
    }
}

inline void build_lane_form(const std::vector<int>& seg, int n_bins, int n_mels, MelPieces& out) {
    constexpr int TL = kTeamLanes;
    const int bpt = out.bpt;
    out.lanes_ok = false;
    out.lane_beg.assign(TL, 0);
    out.lane_seg0.assign(TL, 0);
    out.w4.assign(size_t(TL) * bpt * 4, 0.f);
    out.gtab.assign(size_t(64) * kLaneGatherMax, kLaneZeroRec * kLaneRecWords);
    out.gather_n[0] = out.gather_n[1] = 0;
    if (n_mels > 64 || n_bins + bpt > 2 * (n_bins - 1) || n_bins < 16 + bpt) return;   // (a lane's bpt reads stay inside the spectrum)
    std::vector<LaneRun> runs;
    bool spread = cut_lane_runs(seg, n_bins, bpt, true, runs);
    if (!spread && !cut_lane_runs(seg, n_bins, bpt, false, runs)) return;
    // ---- runs -> lanes.  Spread: the g-th run of a residue class goes to lane group g; idle lanes take the residues a group
    // still misses (they read bins r .. r + bpt - 1 with zero weights).
    std::vector<int> lane_of(runs.size(), -1);
    std::vector<int> run_at(TL, -1);
    if (spread) {
        int seen[16] = {0};
        int fill[4] = {0, 0, 0, 0};
        for (size_t i = 0; i < runs.size(); ++i) {
            const int g = seen[runs[i].beg & 15]++;
            lane_of[i] = 16 * g + fill[g]++;
            run_at[lane_of[i]] = int(i);
        }
        for (int g = 0; g < 4; ++g) {
            bool have[16] = {false};
            for (int l = 16 * g; l < 16 * g + fill[g]; ++l) have[runs[run_at[l]].beg & 15] = true;
            int l = 16 * g + fill[g];
            for (int r = 0; r < 16; ++r) if (!have[r]) out.lane_beg[l++] = r;
        }
    } else {
        for (size_t i = 0; i < runs.size(); ++i) { lane_of[i] = int(i); run_at[i] = int(i); }
    }
    // ---- contributors of every filter: (run, j)
    struct Entry { int run, j; };
    std::vector<std::vector<Entry>> users(n_mels);
    for (size_t i = 0; i < runs.size(); ++i) {
        bool feeds[4] = {false, false, false, false};
        for (int t = 0; t < runs[i].n; ++t) {
            const int sg = seg[runs[i].beg + t];
            if (sg < 0) continue;
            const int r = sg - runs[i].s0;
            if (out.w01[2 * size_t(runs[i].beg + t)] != 0.f) feeds[r] = true;
            if (out.w01[2 * size_t(runs[i].beg + t) + 1] != 0.f) feeds[r + 1] = true;
        }
        for (int j = 0; j < 4; ++j)
            if (feeds[j] && runs[i].s0 + j < n_mels) users[runs[i].s0 + j].push_back({int(i), j});
    }
    int gn[2] = {0, 0};
    for (int m = 0; m < n_mels; ++m) {
        if (int(users[m].size()) > kLaneGatherMax) return;
        if (int(users[m].size()) > gn[m >> 5]) gn[m >> 5] = int(users[m].size());
    }
    // ---- gather order + lane positions: deterministic local search on the modelled wavefront count of the gather (32 filters
    // of a warp read entry t of their lists at once; absent entries all read the zero record)
    std::vector<std::vector<int>> order(n_mels);          // order[m][t] = index into users[m], or -1
    for (int m = 0; m < n_mels; ++m) {
        order[m].assign(gn[m >> 5], -1);
        for (size_t t = 0; t < users[m].size(); ++t) order[m][t] = int(t);
    }
    auto step_cost = [&](int w, int t) {                  // wavefronts of step t of team warp w
        int addr_of_bank[32][8]; int n_in_bank[32] = {0};
        int worst = 1;
        for (int m = 32 * w; m < 32 * w + 32 && m < n_mels; ++m) {
            const int e = order[m][t];
            if (e < 0) continue;
            const int a = lane_of[users[m][e].run] * kLaneRecWords + users[m][e].j, b = a & 31;
            bool dup = false;
            for (int q = 0; q < n_in_bank[b]; ++q) dup = dup || addr_of_bank[b][q] == a;
            if (!dup && n_in_bank[b] < 8) { addr_of_bank[b][n_in_bank[b]++] = a; if (n_in_bank[b] > worst) worst = n_in_bank[b]; }
        }
        return worst;
    };
    auto total_cost = [&]() {
        int c = 0;
        for (int w = 0; w < 2; ++w) for (int t = 0; t < gn[w]; ++t) c += step_cost(w, t);
        return c;
    };
    int cost = total_cost();
    const int ideal = gn[0] + gn[1];
    unsigned long long rng = 0x9E3779B97F4A7C15ull;
    auto next = [&](int n) { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return int((rng >> 33) % unsigned(n)); };
    for (int it = 0; it < 60000 && cost > ideal; ++it) {
        const int kind = next(3);
        if (kind == 0) {                                   // swap two list positions of one filter
            const int m = next(n_mels), n = int(order[m].size());
            if (n < 2) continue;
            const int t0 = next(n), t1 = next(n);
            if (t0 == t1) continue;
            std::swap(order[m][t0], order[m][t1]);
            const int c = total_cost();
            if (c <= cost) cost = c; else std::swap(order[m][t0], order[m][t1]);
        } else {                                           // swap the lanes of two runs: inside a group, or same residue across groups
            const int l0 = next(TL);
            int l1;
            if (kind == 1 || !spread) { l1 = spread ? (l0 & ~15) + next(16) : next(TL); }
            else {
                const int r = (run_at[l0] >= 0 ? runs[run_at[l0]].beg : out.lane_beg[l0]) & 15;
                const int g = next(4);
                l1 = -1;
                for (int l = 16 * g; l < 16 * g + 16; ++l)
                    if (((run_at[l] >= 0 ? runs[run_at[l]].beg : out.lane_beg[l]) & 15) == r) l1 = l;
                if (l1 < 0) continue;
            }
            if (l0 == l1) continue;
            auto swap_lanes = [&]() {
                std::swap(run_at[l0], run_at[l1]);
                std::swap(out.lane_beg[l0], out.lane_beg[l1]);
                if (run_at[l0] >= 0) lane_of[run_at[l0]] = l0;
                if (run_at[l1] >= 0) lane_of[run_at[l1]] = l1;
            };
            swap_lanes();
            const int c = total_cost();
            if (c <= cost) cost = c; else swap_lanes();
        }
    }
    // ---- tables
    for (size_t i = 0; i < runs.size(); ++i) {
        const int lane = lane_of[i];
        out.lane_beg[lane] = runs[i].beg;
        out.lane_seg0[lane] = runs[i].s0;
        for (int t = 0; t < runs[i].n; ++t) {
            const int kk = runs[i].beg + t, sg = seg[kk];
            if (sg < 0) continue;
            const int r = sg - runs[i].s0;                 // 0, 1, 2
            float* w = &out.w4[(size_t(lane) * bpt + t) * 4];
            w[r] = out.w01[2 * size_t(kk)];
            w[r + 1] = out.w01[2 * size_t(kk) + 1];
        }
    }
    for (int m = 0; m < n_mels; ++m)
        for (size_t t = 0; t < order[m].size(); ++t)
            if (order[m][t] >= 0) {
                const Entry& e = users[m][order[m][t]];
                out.gtab[size_t(m) * kLaneGatherMax + t] = lane_of[e.run] * kLaneRecWords + e.j;
            }
    out.gather_n[0] = gn[0];
    out.gather_n[1] = gn[1];
    out.lane_gather_wavefronts = cost;
    out.lane_spread = spread;
    out.lanes_ok = true;
}

// returns "" on success, else an error message
inline std::string build_mel_pieces(const float* fb, int n_bins, int n_mels, MelPieces& out) {
    constexpr int TL = kTeamLanes;
    const int bpt = (n_bins + TL - 1) / TL;
    if (bpt > 64) return "too many bins per lane";
    out.bpt = bpt;
    out.w01.assign(size_t(TL) * bpt * 2, 0.f);
    out.endmask.assign(TL, 0ull);
    out.slot0.assign(TL, 0);
    out.slot1.assign(TL, 0);
    out.pb.assign(n_mels + 2, 0);
    std::vector<int> seg(n_bins, -1);
    for (int k = 0; k < n_bins; ++k) {
        int first = -1, last = -1, count = 0;
        for (int m = 0; m < n_mels; ++m)
            if (fb[size_t(k) * n_mels + m] != 0.f) { if (first < 0) first = m; last = m; ++count; }
        if (count == 0) continue;
        if (count > 2 || last - first > 1) return "mel filterbank row has more than two / non-adjacent non-zeros";
        seg[k] = first;
        out.w01[2 * size_t(k)] = 0.25f * fb[size_t(k) * n_mels + first];
        if (count == 2) out.w01[2 * size_t(k) + 1] = 0.25f * fb[size_t(k) * n_mels + last];
    }
    std::vector<int> piece_seg;
    std::vector<int> first_piece(TL + 1, 0);
    int prev_seg = -1;
    for (int l = 0; l < TL; ++l) {
        first_piece[l] = int(piece_seg.size());
        int cur = -2, last_i = -1;
        for (int i = 0; i < bpt; ++i) {
            const int k = l * bpt + i;
            if (k >= n_bins || seg[k] < 0) continue;
            if (seg[k] < prev_seg) return "mel filterbank centres are not increasing";
            prev_seg = seg[k];
            if (cur != -2 && seg[k] != cur) {                  // the previous contributing bin closed a piece
                out.endmask[l] |= 1ull << last_i;
                piece_seg.push_back(cur);
            }
            cur = seg[k];
            last_i = i;
        }
        if (cur != -2) {
            out.endmask[l] |= 1ull << last_i;
            piece_seg.push_back(cur);
        }
    }
    first_piece[TL] = int(piece_seg.size());
    out.n_pieces = int(piece_seg.size());
    out.max_pieces_per_seg = 0;
    for (size_t i = 0, run = 0; i < piece_seg.size(); ++i) {
        run = (i > 0 && piece_seg[i] == piece_seg[i - 1]) ? run + 1 : 1;
        if (int(run) > out.max_pieces_per_seg) out.max_pieces_per_seg = int(run);
    }
    // pb[j] = first piece with seg >= j - 1   (pieces are sorted by seg)
    for (int j = 0; j < n_mels + 2; ++j) {
        int p = 0;
        while (p < out.n_pieces && piece_seg[p] < j - 1) ++p;
        out.pb[j] = p;
    }
    // record slots
    out.seg_major = n_mels <= 64 && out.max_pieces_per_seg <= kSegMajorRanks;
    for (int l = 0; l < TL && out.seg_major; ++l)
        for (int p = first_piece[l] + 1; p < first_piece[l + 1]; ++p)
            if (piece_seg[p] != piece_seg[p - 1] + 1) out.seg_major = false;          // an empty segment inside the run
    // slot of every piece in the segment-major layout
    std::vector<int> slot_of(piece_seg.size(), 0);
    out.ov.assign(64, kSegMajorZero | (kSegMajorZero << 16));
    int n_over = 0;
    for (size_t p = 0; p < piece_seg.size() && out.seg_major; ++p) {
        int rank = 0;
        while (int(p) - rank - 1 >= 0 && piece_seg[p - rank - 1] == piece_seg[p]) ++rank;
        const int s = piece_seg[p];
        if (rank < 2) {
            slot_of[p] = s + kSegMajorPitch * rank;
        } else {
            slot_of[p] = 2 * kSegMajorPitch + n_over++;
            if (rank == 2) out.ov[s] = (out.ov[s] & ~0xffff) | slot_of[p];
            else out.ov[s] = (out.ov[s] & 0xffff) | (slot_of[p] << 16);
        }
    }
    for (int l = 0; l < TL; ++l) {
        const int p = first_piece[l];
        if (!out.seg_major || p >= first_piece[l + 1]) {
            out.slot0[l] = out.seg_major ? 0 : p;             // (a lane without pieces never stores)
            out.slot1[l] = out.slot0[l] + 1;
            continue;
        }
        out.slot0[l] = slot_of[p];
        out.slot1[l] = piece_seg[p] + 1;                      // later pieces of the lane open their segment: rank 0
    }
    out.n_slots = out.seg_major ? 2 * kSegMajorPitch + n_over : out.n_pieces;

    // ---- lane form
    build_lane_form(seg, n_bins, n_mels, out);
    return "";
}

}  // namespace seld
