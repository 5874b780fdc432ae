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
    std::vector<int> piece0;                      // [64]
    std::vector<int> pb;                          // [n_mels + 2]
    int n_pieces = 0;
    int max_pieces_per_seg = 0;
};

// returns "" on success, else an error message
inline std::string build_mel_pieces(const float* fb, int n_bins, int n_mels, MelPieces& out) {
    constexpr int TL = kTeamLanes;
    const int bpt = (n_bins + TL - 1) / TL;
    if (bpt > 64) return "too many bins per lane";
    out.bpt = bpt;
    out.w01.assign(size_t(TL) * bpt * 2, 0.f);
    out.endmask.assign(TL, 0ull);
    out.piece0.assign(TL, 0);
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
    int prev_seg = -1;
    for (int l = 0; l < TL; ++l) {
        out.piece0[l] = int(piece_seg.size());
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
    return "";
}

}  // namespace seld
