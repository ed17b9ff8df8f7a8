// encode_internal.cuh — host-side interfaces between api.cu and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace dns {

// byte offsets of the Chameleon encode scratch arrays inside one workspace allocation
struct ChamLayout {
    size_t status, sigw, copymap, copymap2, seg_state, incb, tile_bytes, tile_local, group_total, group_off, unres, unres_count, final_tab, carry, total;
};

size_t cham_workspace_bytes(size_t nbytes, int nruns_max, ChamLayout* L);
uint32_t cham_pick_runs(size_t nbytes, int num_sms);
extern int g_cham_decode_impl;  // which Chameleon decode pass kernel runs (chameleon_decode.cu)
extern int g_cham_flag_impl;   // which flag pass kernel runs (chameleon_encode.cu; timing comparisons and tests)
cudaError_t cham_encode_phase1(const uint8_t* d_in, size_t nbytes, uint8_t* ws, const ChamLayout& L, uint32_t nruns,
                               uint32_t* d_table_out, cudaStream_t stream, uint64_t* launches, cudaEvent_t* ev = nullptr);
cudaError_t cham_encode_phase2(const uint8_t* d_in, size_t nbytes, uint8_t* ws, const ChamLayout& L, uint32_t nruns,
                               const uint32_t* d_carry_in, uint8_t* d_out, size_t cap, uint64_t* d_out_size,
                               bool allow_protected_fallback, bool assume_prev_inc, cudaStream_t stream, uint64_t* launches,
                               cudaEvent_t* ev = nullptr);
cudaError_t cham_encode_phase2_blocking(const uint8_t* d_in, size_t nbytes, uint8_t* ws, const ChamLayout& L, uint32_t nruns, uint8_t* d_out,
                                        size_t cap, uint64_t* d_out_size, int max_batches, cudaStream_t stream, uint64_t* launches);
cudaError_t cham_encode_protected_only(const uint8_t* d_in, size_t nbytes, uint8_t* ws, const ChamLayout& L, uint8_t* d_out,
                                       size_t cap, uint64_t* d_out_size, cudaStream_t stream, uint64_t* launches);

// streaming continuation (a reused Codec instance, codec.rs:16,72): the dictionary as the reference keeps it (65536 quads) <-> the
// touched | fingerprint form of the run-parallel encoder; and phase 2 with a carried-in dictionary that gives up (no emit, *ok = false)
// instead of walking in order when the copy map does not settle
cudaError_t cham_quads_to_table(const uint32_t* d_quads, uint32_t* d_table, cudaStream_t stream, uint64_t* launches);
cudaError_t cham_table_into_quads(const uint32_t* d_table, uint32_t* d_quads, cudaStream_t stream, uint64_t* launches);
cudaError_t cham_encode_phase2_stream(const uint8_t* d_in, size_t nbytes, uint8_t* ws, const ChamLayout& L, uint32_t nruns, const uint32_t* d_carry_in,
                                      uint8_t* d_out, size_t cap, uint64_t* d_out_size, uint32_t* d_table_out, int max_batches, cudaStream_t stream,
                                      uint64_t* launches, bool* ok);

// shared pieces of the encoders (chameleon_encode.cu)
struct Status;
size_t prot_state_bytes(uint64_t nseg_max);
cudaError_t prot_debug_read(unsigned long long* out32);   // diagnostics: per fixed-point round {first changed block, changed blocks}   // segment states + candidate tables of prot_iterate for up to nseg_max segments of 256 blocks
cudaError_t prot_iterate_launch(const uint32_t* sigw_or_null, uint64_t nbytes, uint64_t nblocks, uint32_t nseg, Status* st, int it, uint8_t* inc,
                                uint8_t* cm_old, uint8_t* cm_new, uint32_t* in_state, uint32_t* out_state, int block_bytes, int num_sms,
                                cudaStream_t stream);
cudaError_t scan_tiles_launch(const uint32_t* tile_bytes, uint32_t ntiles, uint32_t* tile_local, uint64_t* group_total, uint64_t* group_off,
                              uint32_t ngroups, Status* st, uint64_t cap, uint64_t* d_out_size, cudaStream_t stream);

// cheetah_encode.cu
extern int g_chee_stage_rounds;
size_t chee_workspace_bytes(size_t nbytes, int num_sms);
size_t chee_tables_bytes(int alg, int region, size_t nbytes, int num_sms);
cudaError_t chee_encode_parallel(int alg, const uint8_t* d_in, size_t nbytes, uint8_t* d_out, size_t cap, uint8_t* ws, uint8_t* const tables[3],
                                 uint32_t epoch_base, int num_sms, uint64_t* d_out_size, uint32_t* d_converged, bool resume,
                                 cudaStream_t stream, uint64_t* launches);

// chameleon_decode.cu
size_t cham_decode_workspace_bytes(size_t nbytes, size_t cap, int nruns_max);
cudaError_t cham_decode_parallel(const uint8_t* d_in, size_t nbytes, uint8_t* d_out, size_t cap, uint8_t* ws, int num_sms,
                                 uint64_t* d_out_size, uint32_t* d_nonquiet, cudaStream_t stream, uint64_t* launches);

// cl_decode.cu (run-parallel Cheetah decode)
size_t chee_decode_workspace_bytes(size_t nbytes, size_t cap, int num_sms);
size_t chee_decode_tables_bytes(size_t nbytes, int num_sms);
cudaError_t chee_decode_parallel(const uint8_t* d_in, size_t nbytes, uint8_t* d_out, size_t cap, uint8_t* ws, uint8_t* tables, uint8_t* tail_ws,
                                 int num_sms, uint64_t* d_out_size, uint32_t* d_fallback, cudaStream_t stream, uint64_t* launches);
const void* chee_decode_status_ptr(uint8_t* ws, size_t nbytes, size_t cap, int num_sms, const void** cl_status);

// scalar_codec.cu (Cheetah / Lion, in-order)
size_t scalar_workspace_bytes(int alg);
cudaError_t scalar_encode(int alg, const uint8_t* d_in, size_t nbytes, uint8_t* d_out, size_t cap, uint8_t* ws,
                          uint64_t* d_out_size, cudaStream_t stream, uint64_t* launches, const uint32_t* d_run_if_zero = nullptr, bool keep_state = false);
cudaError_t scalar_decode(int alg, const uint8_t* d_in, size_t nbytes, uint8_t* d_out, size_t cap, uint8_t* ws,
                          uint64_t* d_out_size, cudaStream_t stream, uint64_t* launches, const uint32_t* d_run_if = nullptr, bool keep_state = false);
// tail loop only (codec.rs:102-123), continuing from the state the parallel decoder left: tables already in `ws`, boundary status
// (bounds::DecStatus: tail offset, block count, protection state) and the last hash (cheedec::ClStatus::final_ctx) on the device;
// runs only if *d_skip_if == 0
cudaError_t scalar_decode_tail(int alg, const uint8_t* d_in, size_t nbytes, uint8_t* d_out, size_t cap, uint8_t* ws, const void* d_bounds_status,
                               const void* d_cl_status, uint64_t* d_out_size, cudaStream_t stream, uint64_t* launches, const uint32_t* d_skip_if);

// table helpers (sharded API, pipelined host path)
cudaError_t cham_status_accumulate(const uint8_t* ws, const ChamLayout& L, uint32_t* d_flag, cudaStream_t stream, uint64_t* launches);
cudaError_t cham_table_init(uint32_t* d_table, cudaStream_t stream, uint64_t* launches);
cudaError_t cham_rank_fold(const uint32_t* d_tables, uint32_t rank, uint32_t* d_carry, cudaStream_t stream, uint64_t* launches);
cudaError_t cham_seam_words(const uint8_t* ws, const ChamLayout& L, size_t nbytes, const uint64_t* d_out_size, uint32_t* d_words, cudaStream_t stream, uint64_t* launches);
cudaError_t cham_seam_verdict(const uint32_t* d_all_words, uint32_t world, uint32_t rank, uint32_t* d_flags, uint64_t* d_total, uint64_t* d_offsets,
                              cudaStream_t stream, uint64_t* launches);
cudaError_t cham_table_fold(uint32_t* d_acc, const uint32_t* d_next, cudaStream_t stream, uint64_t* launches);

}  // namespace dns
