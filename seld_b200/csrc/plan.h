// Internal plan object shared by the .cu translation units (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/seld_b200.h"

struct seld_plan {
    int sample_rate, n_fft, win_length, hop, n_mels, n_chan, mode;
    int tf_variant;  // created with SELD_MODE_FOA_TF: magnitude mel, 20 log10 without a floor, zero-padded tail frames
    int n_bins;      // n_fft / 2 + 1
    int n_out_ch;    // 7 | 10
    int device;
    int num_sms;
    int max_smem_optin;
    // device tables
    float* window;   // [n_fft]
    float* twiddle;  // [n_fft][2]  exp(-2 pi i j / n_fft)
    float* tw_t;     // [n_fft/32][32][2]  tw_t[k2][lane] = exp(-2 pi i lane*k2 / n_fft)
    float* w01;      // [64 * bins_per_lane][2]  piece form of the mel bank (mel_pieces.h)
    unsigned long long* endmask;  // [64]
    int* slot0;      // [64]  record slot of each team lane's first / second piece (mel_pieces.h)
    int* slot1;      // [64]
    int* ov;         // [64]  overflow slots of the segment-major record layout
    int* pb;         // [n_mels + 2]
    float* w4;       // [64 * bins_per_lane][4]  flush-free lane form of the bank (mel_pieces.h)
    int* lane_beg;   // [64]
    int* gtab;       // [64][kLaneGatherMax]
    int gather_n[2];
    int lanes_ok;
    void* gcc_bt;    // MIC, n_fft 1024, 64 lags: fp16 [64][1024] basis of the tensor-core lag projection (else null)
    int n_pieces;
    int n_slots;     // piece records per frame in the layout in use
    int seg_major;   // segment-major record layout (fast gather) instead of the compact one
    int max_pieces_per_seg;
    // extract launch geometry (warps and shared memory are chosen per kernel variant at launch, extract.cu)
    int grid;
    // stats launch geometry
    int stats_blocks;
};

namespace seld {
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
// one more kernel of this library was launched (seld_launch_count: what bench.py reports as gpu_launches)
void note_launch();
// multiprocessor count of the CURRENT device, queried once per device (a process may drive several GPUs)
int device_sm_count();
// true the first time it is called for (slot, current device): "set this kernel attribute once per device"
bool first_use_on_device(unsigned long long* slot_bits);
}  // namespace seld

#define SELD_CUDA_TRY(expr)                                          \
    do {                                                             \
        cudaError_t _e = (expr);                                     \
        if (_e != cudaSuccess) return seld::cuda_fail(_e, #expr);    \
    } while (0)
