// common.cuh -- host-side state shared by the translation units of libg2048.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/g2048.h"

namespace g2048 {

constexpr int kMaxDevices = 32;
constexpr int kRowEntries = 65536;
constexpr size_t kRowTableBytes = kRowEntries * sizeof(uint16_t);    // 128 KiB
constexpr size_t kCodeTableBytes = kRowEntries * sizeof(uint8_t);    //  64 KiB

struct DeviceState {
    bool ready = false;
    uint16_t *row = nullptr;                 // LEFT-move result per 16-bit row
    uint8_t *code = nullptr;                 // merge codes per 16-bit row
    unsigned long long *overflow = nullptr;  // sticky nibble-saturation counter
    int sm_count = 0;
    // Stream-ordered pool for per-launch scratch.  Its release threshold is unlimited: the default pool hands
    // free memory back to the driver at every synchronisation, so a host call that ends in a stream sync would
    // pay a fresh device allocation (~0.4 ms) on its next launch.
    cudaMemPool_t pool = nullptr;
    // 512-entry pair table (board.cuh: pair_table_entry) followed by 256 float2 observation pairs (entry b = the
    // float32 values e / 15 of the two nibbles of byte b): the per-step kernel copies them into shared memory
    // instead of computing them in every block
    uint32_t *step_tables = nullptr;
};
constexpr int kStepTableWords = 512 + 2 * 256;

// Returns the state of the current device, or nullptr (and sets the error) if g2048_init
// has not been called for it.
DeviceState *current_device_state();
int set_error(int code, const char *fmt, ...);
int check_cuda(cudaError_t e, const char *what);
void count_launch(int n = 1);

#define G2048_CUDA(call)                                          \
    do {                                                          \
        int _rc = ::g2048::check_cuda((call), #call);             \
        if (_rc != G2048_OK) return _rc;                          \
    } while (0)

// Launch helpers implemented in beam.cu
int set_tuning(int key, int value);
int launch_beam_search(DeviceState *st, const uint64_t *roots, const uint8_t *legal, const uint32_t *call,
                       uint32_t call0, uint8_t *action, float *prob, double *best_score, int32_t *nodes,
                       int64_t n, int beam_width, int search_depth, int early_thr, int mid_thr,
                       uint64_t seed, uint32_t game0, cudaStream_t stream);
int launch_play_games(DeviceState *st, int64_t n, int beam_width, int search_depth, int early_thr, int mid_thr,
                      int max_moves, uint64_t seed, uint32_t game0, int32_t *score, uint8_t *highest_exp,
                      int32_t *moves, int32_t *valid, int32_t *invalid, int32_t *milestone, int64_t *nodes,
                      uint64_t *final_board, cudaStream_t stream);

}  // namespace g2048
