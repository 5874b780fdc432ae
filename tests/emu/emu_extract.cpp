// CPU emulation of one frame team (two warps) of the fused extractor: runs the per-lane phases of
// seld_b200/csrc/extract_core.cuh lane by lane (a phase boundary is a __syncwarp() on the device).
// TEST INFRASTRUCTURE: lets the CPU suite check the kernel's index math, tables and FFT against the
// oracle without a GPU.  Built by tests/test_emu_cpu.py with g++ -std=c++17.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../seld_b200/csrc/extract_core.cuh"
#include "../../seld_b200/csrc/mel_pieces.h"

using namespace seld;

template <int R, int MODE, int LAYOUT>
static void run(const float* wav, int n_clips, long long L, int hop, int n_mels, const Tables& tb, bool fast_gather,
                int T_out, float* out, float* clip_max) {
    using G = Geo<R>;
    constexpr int C = (MODE == MODE_FOA) ? 7 : 10;
    const int layout = LAYOUT;
    float wreg[32][R];
    for (int l = 0; l < 32; ++l) for (int n2 = 0; n2 < R; ++n2) wreg[l][n2] = tb.window[l + 32 * n2];
    const int T_raw = 1 + int(L / hop);
    // team buffers as on the device: one exchange buffer per warp (the spectrum overwrites it in place), piece buffer X
    std::vector<float2> EA(G::E_ELEMS), EB(G::E_ELEMS), X(G::E_ELEMS + 4096);
    std::vector<float2> cols(size_t(32) * G::COLS * 32);
    std::vector<float> acc(size_t(n_mels) * C, 0.f);
    auto stage2_inplace = [&](std::vector<float2>& buf) {
        for (int l = 0; l < 32; ++l) stage2_load_fft<R>(buf.data(), cols.data() + size_t(l) * G::COLS * 32, l);
        for (int l = 0; l < 32; ++l) stage2_store<R>(cols.data() + size_t(l) * G::COLS * 32, buf.data(), l);
    };
    for (int clip = 0; clip < n_clips; ++clip) {
        ClipSrc src;
        src.base = wav + size_t(clip) * 4 * L;
        src.n_samples = L;
        if (layout == LAYOUT_PLANAR_CL) { src.chan_stride = L; src.samp_stride = 1; }
        else { src.chan_stride = 1; src.samp_stride = 4; }
        float cmax = -INFINITY;
        const int T_tot = T_raw > T_out ? T_raw : T_out;
        for (int t = 0; t < T_tot; ++t) {
            float* row = (t < T_out) ? out + (size_t(clip) * T_out + t) * n_mels * C : nullptr;
            if (t >= T_raw) { memset(row, 0, sizeof(float) * n_mels * C); continue; }
            const long long start = (long long)t * hop - G::N / 2;
            for (int l = 0; l < 32; ++l) stage1_forward<R, LAYOUT>(src, 0, 1, start, wreg[l], tb, EA.data(), l);
            stage2_inplace(EA);
            for (int l = 0; l < 32; ++l) stage1_forward<R, LAYOUT>(src, 2, 3, start, wreg[l], tb, EB.data(), l);
            stage2_inplace(EB);
            if (MODE == MODE_FOA && R == 32 && tb.w4) {                      // the flush-free lane form, as the device selects it
                for (int u = 0; u < G::TL; ++u) bin_phase_lanes<R>(EA.data(), EB.data(), tb, X.data(), 1e-8f, u);
                for (int u = 0; u < G::TL; ++u) cmax = fmaxf(cmax, gather_records<>(X.data(), tb, acc.data(), n_mels, u));
            } else {
                for (int u = 0; u < G::TL; ++u) bin_phase<R, MODE>(EA.data(), EB.data(), tb, X.data(), 1e-8f, u);
                for (int u = 0; u < G::TL; ++u)
                    cmax = fmaxf(cmax, fast_gather ? gather_lanes<MODE>(X.data(), tb, acc.data(), n_mels, u)
                                                   : gather_phase<MODE>(X.data(), tb, acc.data(), n_mels, u));
            }
            if (MODE == MODE_MIC) {
                for (int l = 0; l < 32; ++l) gcc_stage1<R, 0>(EA.data(), EB.data(), X.data(), l);
                for (int l = 0; l < 32; ++l) gcc_stage2<R, 0>(X.data(), tb, acc.data(), n_mels, l);
                for (int l = 0; l < 32; ++l) gcc_stage1<R, 1>(EA.data(), EB.data(), X.data(), l);
                for (int l = 0; l < 32; ++l) gcc_stage2<R, 1>(X.data(), tb, acc.data(), n_mels, l);
                for (int l = 0; l < 32; ++l) gcc_stage1<R, 2>(EA.data(), EB.data(), X.data(), l);
                for (int l = 0; l < 32; ++l) gcc_stage2<R, 2>(X.data(), tb, acc.data(), n_mels, l);
                std::fill(X.begin(), X.end(), make_float2(0.f, 0.f));        // X goes back to being the piece buffer
            }
            if (row) for (int u = 0; u < G::TL; ++u) store_row(acc.data(), n_mels * C, row, u, G::TL);
        }
        clip_max[clip] = cmax;
    }
}

extern "C" int emu_extract(const float* wav, int layout, int n_clips, long long L, int n_fft, int hop, int n_mels,
                           int mode, const float* window, const float* twiddle, const float* mel_fb, int T_out,
                           float* out, float* clip_max) {
    // lane-contiguous twiddle table tw_t[k2*32 + lane] = W^(lane*k2), as seld_plan_create builds it
    const float2* lin = reinterpret_cast<const float2*>(twiddle);
    std::vector<float2> tw_t(n_fft);
    for (int k2 = 0; k2 < n_fft / 32; ++k2)
        for (int l = 0; l < 32; ++l) tw_t[k2 * 32 + l] = lin[(l * k2) % n_fft];
    MelPieces mp;
    if (!build_mel_pieces(mel_fb, n_fft / 2 + 1, n_mels, mp).empty()) return -2;
    Tables tb{window, tw_t.data(), lin, reinterpret_cast<const float2*>(mp.w01.data()), mp.endmask.data(), mp.slot0.data(), mp.slot1.data(),
              mp.pb.data(), mp.ov.data(), mp.lanes_ok ? mp.w4.data() : nullptr, mp.lane_beg.data(), mp.gtab.data(),
              mp.gather_n[0], mp.gather_n[1]};
#define GO3(RR, MM, LL) run<RR, MM, LL>(wav, n_clips, L, hop, n_mels, tb, mp.seg_major, T_out, out, clip_max)
#define GO(RR)                                                                   \
    if (n_fft == 32 * RR) {                                                      \
        if (mode == MODE_FOA) { if (layout == 0) GO3(RR, MODE_FOA, 0); else GO3(RR, MODE_FOA, 1); } \
        else { if (layout == 0) GO3(RR, MODE_MIC, 0); else GO3(RR, MODE_MIC, 1); } \
        return 0;                                                                \
    }
    GO(8) GO(16) GO(32) GO(64)
    return -1;
}

extern "C" float emu_key_roundtrip(float f) { return key_to_float(float_to_key(f)); }

// Mel piece tables as seld_plan_create builds them (for tests/test_emu_cpu.py::test_mel_piece_layout_invariants).
// info: [bpt, n_pieces, max_pieces_per_seg, seg_major, n_slots, zero_slot, pitch]
extern "C" int emu_mel_layout(const float* mel_fb, int n_bins, int n_mels, int* info, int* slot0, int* slot1, int* ov,
                              unsigned long long* endmask) {
    MelPieces mp;
    if (!build_mel_pieces(mel_fb, n_bins, n_mels, mp).empty()) return -2;
    info[0] = mp.bpt; info[1] = mp.n_pieces; info[2] = mp.max_pieces_per_seg; info[3] = mp.seg_major ? 1 : 0;
    info[4] = mp.n_slots; info[5] = kSegMajorZero; info[6] = kSegMajorPitch;
    for (int l = 0; l < kTeamLanes; ++l) { slot0[l] = mp.slot0[l]; slot1[l] = mp.slot1[l]; endmask[l] = mp.endmask[l]; }
    for (int s = 0; s < 64; ++s) ov[s] = mp.ov[s];
    return 0;
}

// Lane form of the bank (for tests/test_emu_cpu.py::test_mel_lane_form_invariants).
// info: [lanes_ok, bpt, gather_n0, gather_n1, rec_words, gather_max, zero_rec, spread, modelled gather wavefronts]
extern "C" int emu_mel_lanes(const float* mel_fb, int n_bins, int n_mels, int* info, int* lane_beg, int* gtab, float* w4) {
    MelPieces mp;
    if (!build_mel_pieces(mel_fb, n_bins, n_mels, mp).empty()) return -2;
    info[0] = mp.lanes_ok ? 1 : 0; info[1] = mp.bpt; info[2] = mp.gather_n[0]; info[3] = mp.gather_n[1];
    info[4] = kLaneRecWords; info[5] = kLaneGatherMax; info[6] = kLaneZeroRec; info[7] = mp.lane_spread ? 1 : 0;
    info[8] = mp.lane_gather_wavefronts;
    for (int l = 0; l < kTeamLanes; ++l) lane_beg[l] = mp.lane_beg[l];
    for (size_t i = 0; i < mp.gtab.size(); ++i) gtab[i] = mp.gtab[i];
    for (size_t i = 0; i < mp.w4.size(); ++i) w4[i] = mp.w4[i];
    return 0;
}
