/*
 * density_b200.h — C ABI of the B200-native (sm_100a) implementation of density's
 * Chameleon / Cheetah / Lion encode/decode hot path.
 *
 * Drop-in boundary. The first nine symbols have exactly the names, signatures and
 * semantics of the reference's own `extern "C"` exports, so a binary (or a Rust
 * `extern "C"` block) that links against density-rs can link
 * against libdensity_b200.so instead:
 *
 *   chameleon_encode / chameleon_decode / chameleon_safe_encode_buffer_size
 *       replace /root/reference/src/algorithms/chameleon/chameleon.rs:70-83
 *   cheetah_encode / cheetah_decode / cheetah_safe_encode_buffer_size
 *       replace /root/reference/src/algorithms/cheetah/cheetah.rs:105-118
 *   lion_encode / lion_decode / lion_safe_encode_buffer_size
 *       replace /root/reference/src/algorithms/lion/lion.rs:193-206
 *
 * They take plain pointers and sizes. The pointers may be HOST pointers (pageable or
 * pinned; the library stages through device memory) or DEVICE pointers (detected with
 * cudaPointerGetAttributes; no staging). The call is synchronous and returns the number
 * of bytes written, or 0 on any error (the reference maps Err -> 0 the same way,
 * chameleon.rs:72 `unwrap_or(0)`; where the reference would panic on an undersized
 * buffer — io/write_buffer.rs:19 — this library returns 0 and never writes out of bounds).
 * Output is bit-identical to the reference's Codec::encode / Codec::decode
 * (codec/codec.rs:72-126) on the same input.
 *
 * There is no CPU fallback: every entry point fails (returns 0 / an error code) when no
 * CUDA device is usable.
 */
#ifndef DENSITY_B200_H
#define DENSITY_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define DENSITY_B200_API __attribute__((visibility("default")))
#else
#define DENSITY_B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* ---- the reference's FFI surface (host or device pointers, synchronous) ---------------- */
DENSITY_B200_API size_t chameleon_encode(const uint8_t* input, size_t input_size, uint8_t* output, size_t output_size);
DENSITY_B200_API size_t chameleon_decode(const uint8_t* input, size_t input_size, uint8_t* output, size_t output_size);
DENSITY_B200_API size_t chameleon_safe_encode_buffer_size(size_t size); /* codec/codec.rs:18-21 */

DENSITY_B200_API size_t cheetah_encode(const uint8_t* input, size_t input_size, uint8_t* output, size_t output_size);
DENSITY_B200_API size_t cheetah_decode(const uint8_t* input, size_t input_size, uint8_t* output, size_t output_size);
DENSITY_B200_API size_t cheetah_safe_encode_buffer_size(size_t size);

DENSITY_B200_API size_t lion_encode(const uint8_t* input, size_t input_size, uint8_t* output, size_t output_size);
/* PERFORMANCE LIMIT: lion_decode runs on the exact in-order device kernel (one thread, ~10-20 MB/s): no parallel formulation of the
   5-deep move-to-front prediction lists that beats in-order is known. Results are bit-exact; for streams
   beyond a few MiB the reference's CPU decoder is the faster choice for this one symbol. All other eight symbols run parallel kernels. */
DENSITY_B200_API size_t lion_decode(const uint8_t* input, size_t input_size, uint8_t* output, size_t output_size);
DENSITY_B200_API size_t lion_safe_encode_buffer_size(size_t size);

/* ---- device-resident, stream-ordered variants (what bench.py times) --------------------- */
#define DENSITY_B200_CHAMELEON 0
#define DENSITY_B200_CHEETAH 1
#define DENSITY_B200_LION 2

#define DENSITY_B200_OK 0
#define DENSITY_B200_ECUDA 1      /* a CUDA runtime call failed (see density_b200_last_error) */
#define DENSITY_B200_ECAPACITY 2  /* output buffer too small */
#define DENSITY_B200_EMALFORMED 3 /* truncated / malformed stream on decode */
#define DENSITY_B200_EARG 4       /* bad argument (alignment, algorithm id, null pointer) */

/*
 * Encode `n` bytes at device pointer d_in (4-byte aligned) into d_out (2-byte aligned,
 * capacity `cap` >= *_safe_encode_buffer_size(n)). All work is enqueued on `stream`
 * (a cudaStream_t passed as void*; NULL = legacy default stream); nothing is synchronised.
 * The encoded size is written to *d_out_size (device memory, 8 bytes) when the stream
 * reaches that point; it is 0 if the device-side capacity check failed.
 * Returns DENSITY_B200_OK or an error code for failures detectable at enqueue time.
 */
DENSITY_B200_API int density_b200_encode_device(int alg, const uint8_t* d_in, size_t n, uint8_t* d_out, size_t cap,
                               uint64_t* d_out_size, void* stream);
/* Test/diagnostic variant: choose the encode path explicitly.
   Chameleon: path 0 = auto (run-parallel fast path, exact protection-aware fallback when needed),
   1 = fast path only (no fallback; out size is only valid if the stream is "quiet"),
   2 = exact in-order protection-aware walk only, 3 = in-order single-thread kernel,
   4 = like 0, but the call may BLOCK on the stream: if the copy map has not settled after the 5 rounds that are always enqueued,
   the host keeps iterating (up to 96 more rounds) before the in-order walk takes over; this is what the nine reference
   symbols use (they are synchronous anyway).
   Cheetah / Lion: 0 = auto (run-parallel encoder; the in-order kernel, queued behind it, runs only if the copy map did
   not settle), 1 = run-parallel encoder only (*d_out_size == 0 if the copy map did not settle), 3 = in-order kernel,
   4 = like 0 but may BLOCK: the host reads the verdict and resumes the iteration (up to 12 times) before the in-order kernel. */
DENSITY_B200_API int density_b200_encode_device_path(int alg, const uint8_t* d_in, size_t n, uint8_t* d_out, size_t cap,
                                    uint64_t* d_out_size, void* stream, int path);
/* Decode counterpart: path 0 = auto (parallel Chameleon decoder, also for streams with copy-mode blocks; the exact in-order
   kernel is queued behind it as a safety net and for Cheetah / Lion), 1 = parallel decoder only (size 0 if it had to give
   up), 3 = in-order kernel only. */
DENSITY_B200_API int density_b200_decode_device_path(int alg, const uint8_t* d_in, size_t n, uint8_t* d_out, size_t cap,
                                    uint64_t* d_out_size, void* stream, int path);
/* Diagnostic: status of the last Chameleon encode on the current device (synchronises the device):
   out6 = {out_bytes, nonquiet (copy mode was needed), error, first_nonquiet_block, copy map converged, scratch}. */
DENSITY_B200_API int density_b200_encode_status(uint64_t* out6);
/* Diagnostic: status of the last parallel Chameleon decode on the current device (synchronises the device):
   out10 = {out_bytes, main_blocks, tail_off, nonquiet, error, last_main_inc, in_order_boundaries, penalty, penalty_start, prev_incompressible}.
   in_order_boundaries != 0: the stream had copy-mode blocks (codec.rs:89-92) and the boundaries came from the in-order walk. */
DENSITY_B200_API int density_b200_decode_status(uint64_t* out10);
/* Diagnostic: the context iteration of the last run-parallel Cheetah decode on the current device (synchronises the device):
   out4 = {rounds used, settled (0: the in-order kernel took over), run walks after round 0, round budget}. */
DENSITY_B200_API int density_b200_cheetah_decode_rounds(uint32_t* out4);
/* Same contract for decode; `cap` must be >= the original length. */
DENSITY_B200_API int density_b200_decode_device(int alg, const uint8_t* d_in, size_t n, uint8_t* d_out, size_t cap,
                               uint64_t* d_out_size, void* stream);

/*
 * Sharded Chameleon encode (one bit-exact stream cut across several GPUs / calls).
 * The stream is cut at multiples of 256 bytes. Every shard runs phase 1 independently, the
 * 256 KiB last-writer tables are exchanged by the caller (torch.distributed all_gather in
 * density_b200/sharded.py), and phase 2 finishes the shard given the dictionary carried in
 * from all earlier shards. The concatenation of the shard outputs equals the output of one
 * chameleon_encode call over the concatenated input, provided the protection automaton stays
 * quiet (reported through *d_flags bit 0 otherwise; the caller then falls back to one device).
 */
typedef struct density_b200_shard density_b200_shard; /* opaque */
DENSITY_B200_API density_b200_shard* density_b200_shard_create(void);
DENSITY_B200_API void density_b200_shard_destroy(density_b200_shard*);
/* phase 1: flags with unknown carry-in; exports this shard's last-writer table (65536 x u32:
   bit 16 = bucket touched, low 16 bits = fingerprint) to d_table_out. */
DENSITY_B200_API int density_b200_shard_phase1(density_b200_shard*, const uint8_t* d_in, size_t n, int is_last_shard,
                              uint32_t* d_table_out, void* stream);
/* phase 2: d_carry_in = table state before this shard (65536 x u32, same encoding; for the first
   shard pass NULL). Writes the shard's piece of the stream to d_out and its size to *d_out_size. */
DENSITY_B200_API int density_b200_shard_phase2(density_b200_shard*, const uint32_t* d_carry_in, uint8_t* d_out, size_t cap,
                              uint64_t* d_out_size, uint32_t* d_flags, void* stream);
/* Fold shard tables left to right: d_acc = (d_next touched) ? d_next : d_acc, elementwise, 65536 entries.
   d_acc == NULL-initialised state is produced by density_b200_table_init. */
DENSITY_B200_API int density_b200_table_init(uint32_t* d_table, void* stream);
DENSITY_B200_API int density_b200_table_fold(uint32_t* d_acc, const uint32_t* d_next, void* stream);

/*
 * The same in C++ end to end (what bench.py --gpus N runs): one process per GPU, the exchange over NCCL (NVLink / NVSwitch).
 *   density_b200_sharded_unique_id   rank 0 makes the 128-byte NCCL id; the caller hands it to every rank (e.g. torch.distributed broadcast)
 *   density_b200_sharded_create      joins the communicator (world == 1: no NCCL needed, id may be NULL)
 *   density_b200_encode_sharded      phase 1 -> ncclAllGather of the 256 KiB tables -> ONE fold kernel -> phase 2 -> seam verdict
 *                                    (ncclAllGather of 32 bytes per rank: first / last block incompressible, quiet, size) -> optional
 *                                    variable-length gather of the pieces to `gather_root` (grouped ncclSend / ncclRecv at prefix-sum
 *                                    offsets). *d_flags != 0: the stream is not quiet (a copy-mode block somewhere, or two incompressible
 *                                    blocks across a cut): the pieces are void and the caller encodes on one device instead.
 *                                    *d_total_size = length of the whole stream, on every rank. gather_root < 0: no gather, nothing blocks.
 */
typedef struct density_b200_sharded density_b200_sharded; /* opaque */
DENSITY_B200_API int density_b200_sharded_unique_id(uint8_t* out128);
DENSITY_B200_API density_b200_sharded* density_b200_sharded_create(const uint8_t* nccl_unique_id_128, int rank, int world);
DENSITY_B200_API void density_b200_sharded_destroy(density_b200_sharded*);
DENSITY_B200_API int density_b200_encode_sharded(density_b200_sharded*, const uint8_t* d_in, size_t n, uint8_t* d_out, size_t cap, uint64_t* d_out_size,
                                uint32_t* d_flags, uint64_t* d_total_size, int gather_root, uint8_t* d_gather, size_t gather_cap, void* stream);
/* stage times (ms) of the last call: [0] flag pass, [1] table exchange + fold, [2] carry / resolve / sizes / scan, [3] emit, [4] seams + gather */
DENSITY_B200_API int density_b200_sharded_profile(density_b200_sharded*, float* out_ms5);

/*
 * A reused Codec INSTANCE (streaming continuation). In the reference `encode` / `decode` are methods of an instance
 * (/root/reference/src/codec/codec.rs:16,72,82) whose dictionary survives from call to call until clear_state()
 * (chameleon.rs:148-150, cheetah.rs:198-202, lion.rs:327-331), while the protection state is created inside every call
 * (codec.rs:75,85); the nine symbols above build a fresh instance per call (chameleon.rs:45-53). These mirror the instance:
 *   density_b200_codec_create(alg)  = X::new()        density_b200_codec_clear_state = Codec::clear_state
 *   density_b200_codec_encode       = Codec::encode   density_b200_codec_decode      = Codec::decode
 * Synchronous, host or device pointers, bytes written or 0 on error. Chameleon encode runs the run-parallel kernels with the
 * instance's dictionary carried in (buffers larger than HBM can be encoded piecewise, bit-exact with an instance on the CPU);
 * Cheetah / Lion and every decode run the exact in-order kernel on the instance's tables.
 */
typedef struct density_b200_codec density_b200_codec; /* opaque */
DENSITY_B200_API density_b200_codec* density_b200_codec_create(int alg);
DENSITY_B200_API void density_b200_codec_destroy(density_b200_codec*);
DENSITY_B200_API int density_b200_codec_clear_state(density_b200_codec*);
DENSITY_B200_API size_t density_b200_codec_encode(density_b200_codec*, const uint8_t* input, size_t input_size, uint8_t* output, size_t output_size);
DENSITY_B200_API size_t density_b200_codec_decode(density_b200_codec*, const uint8_t* input, size_t input_size, uint8_t* output, size_t output_size);

/* ---- per-stage device timing of the last Chameleon encode on the current device ------- */
/* When enabled, density_b200_encode_device records CUDA events on the caller's stream around the flag pass
   and the emit pass of every call (ring of 64 calls; enable(1) resets it). density_b200_profile_get waits
   for them and returns the per-call AVERAGE over the recorded calls:
   out_ms[0] = flag pass, out_ms[1] = carry/resolve/sizes/scan, out_ms[2] = emit (milliseconds). */
DENSITY_B200_API void density_b200_profile_enable(int enable);
DENSITY_B200_API int density_b200_profile_get(float* out_ms);

/* ---- housekeeping ---------------------------------------------------------------------- */
/* Last error message of the calling thread's most recent failing call ("" if none). */
DENSITY_B200_API const char* density_b200_last_error(void);
/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
DENSITY_B200_API uint64_t density_b200_kernel_launches(void);
/* 1 if the last Chameleon encode on this device used the segment-parallel fast path end to end,
   0 if it had to fall back to the sequential protection-aware path. */
DENSITY_B200_API int density_b200_last_encode_was_fast(void);
/* Free all cached device workspaces. */
DENSITY_B200_API void density_b200_shutdown(void);
/* Test hook: cut every stage of the Cheetah / Lion copy-map iteration to k rounds (1..7, default 7) so that the host-resumed
   iteration of path 4 can be exercised on ordinary inputs. */
DENSITY_B200_API void density_b200_test_set_stage_rounds(int k);
/* Test / timing hook: which Chameleon flag pass kernel runs (1 = round-1 class protocol, 6 = write / verify / replay; default 6). */
DENSITY_B200_API void density_b200_test_set_flag_impl(int k);
/* Same for the Chameleon decode pass (1 = round-1 kernel, 7 = write / verify / mailbox; default 7). */
DENSITY_B200_API void density_b200_test_set_decode_impl(int k);
/* Diagnostic: the last copy-map iteration on the current device, per fixed-point round {first block whose copy status changed
   (~0: none), number of such blocks}: 16 rounds x 2 values. Synchronises the device. */
DENSITY_B200_API int density_b200_prot_debug(uint64_t* out32);
/* Library version string. */
DENSITY_B200_API const char* density_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* DENSITY_B200_H */
