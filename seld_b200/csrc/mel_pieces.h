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
};

constexpr int kLaneSegs = 3;
constexpr int kLaneRecWords = 30;                 // 7 channels x (F0, F1, F2, F3) = 28 words, padded to an even, bank-friendly pitch
constexpr int kLaneGatherMax = 8;
constexpr int kLaneZeroRec = 64;                  // record 64 is never written: absent table entries point at it

constexpr int kSegMajorRanks = 4;
constexpr int kSegMajorPitch = 65;      // odd, so the pieces of one segment land in different bank groups
constexpr int kSegMajorZero = 64 + kSegMajorPitch;      // slot 129 = rank 1 of the non-existent segment 64

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

    // ---- lane form: greedy cut of the bin axis into <= 64 runs of <= bpt bins over <= kLaneSegs consecutive segments
    out.lanes_ok = false;
    out.lane_beg.assign(TL, 0);
    out.lane_seg0.assign(TL, 0);
    out.w4.assign(size_t(TL) * bpt * 4, 0.f);
    out.gtab.assign(size_t(64) * kLaneGatherMax, kLaneZeroRec * kLaneRecWords);
    out.gather_n[0] = out.gather_n[1] = 0;
    if (n_mels <= 64) {
        bool ok = true;
        int lane = 0, k = 0;
        std::vector<std::vector<int>> users(n_mels);          // filter -> (lane * kLaneRecWords + j) of every record that feeds it
        while (k < n_bins && ok) {
            if (lane >= TL) { ok = false; break; }
            const int beg = k;
            int s0 = -1, s_last = -1, n = 0;
            while (k < n_bins && n < bpt) {
                const int sg = seg[k];
                if (sg >= 0) {
                    if (s0 < 0) s0 = sg;
                    if (sg - s0 >= kLaneSegs) break;           // a fourth segment: the next lane takes it
                    s_last = sg;
                }
                ++k; ++n;
            }
            out.lane_beg[lane] = beg;
            out.lane_seg0[lane] = s0 < 0 ? 0 : s0;
            bool feeds[4] = {false, false, false, false};
            for (int i = 0; i < k - beg; ++i) {
                const int kk = beg + i, sg = seg[kk];
                if (sg < 0) continue;
                const int r = sg - s0;                         // 0, 1, 2
                const float w0 = out.w01[2 * size_t(kk)], w1 = out.w01[2 * size_t(kk) + 1];
                float* w = &out.w4[(size_t(lane) * bpt + i) * 4];
                if (r == 0) { w[0] = w0; w[1] = w1; }
                else if (r == 1) { w[1] = w0; w[2] = w1; }
                else { w[2] = w0; w[3] = w1; }
                if (w0 != 0.f) feeds[r] = true;
                if (w1 != 0.f) feeds[r + 1] = true;
            }
            for (int j = 0; j < 4; ++j)
                if (feeds[j] && s0 + j < n_mels) users[s0 + j].push_back(lane * kLaneRecWords + j);
            ++lane;
        }
        for (int m = 0; m < n_mels && ok; ++m) {
            if (int(users[m].size()) > kLaneGatherMax) { ok = false; break; }
            for (size_t t = 0; t < users[m].size(); ++t) out.gtab[size_t(m) * kLaneGatherMax + t] = users[m][t];
            int& gn = out.gather_n[m < 32 ? 0 : 1];
            if (int(users[m].size()) > gn) gn = int(users[m].size());
        }
        // lanes past the last run keep beg = 0 and zero weights: they read bins 0 .. bpt - 1 and contribute nothing
        out.lanes_ok = ok && (n_bins + bpt <= 2 * (n_bins - 1));   // a lane's bpt reads stay inside the n_fft-long spectrum buffer
    }
    return "";
}

}  // namespace seld
