/*
 * ertgpu.h -- C ABI of libertgpu.so: a Blackwell (sm_100a) implementation of the
 * rtlamr `protocol.Decoder` hot path (uint8 IQ -> magnitude -> Manchester
 * matched filter -> quantize -> pack -> preamble search -> slice -> CRC screen).
 *
 * This is the drop-in boundary.  A thin cgo shim (go/protocol/decode_cuda.go,
 * shown in INTEGRATION.md) keeps the reference's Go names and calls these
 * entry points instead of running decode.go's loops.  Every entry point cites
 * the reference interface it replaces (paths relative to the rtlamr checkout).
 *
 * Conventions
 *  - plain C types only; no C++/torch types cross the boundary; nothing throws.
 *  - every function returns 0 (ERTGPU_OK) or a negative ERTGPU_E* code;
 *    ertgpu_last_error(h) gives the message.  The Go shim turns errors into
 *    panics where the reference panics (short input: decode.go:222).
 *  - one handle == one protocol.Decoder == one sample stream.  Calls on a handle
 *    must be serialised by the caller (the reference's Decode is not re-entrant
 *    either: it is called from one goroutine, main.go:235).  Different handles
 *    are independent; any OS thread may call (the library sets the device per
 *    call and keeps no thread-local state).
 *  - the library never keeps a caller pointer after a call returns (cgo rule).
 *  - there is NO CPU fallback: without a usable CUDA device every call fails
 *    with ERTGPU_ECUDA.
 */
#ifndef ERTGPU_H
#define ERTGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ERTGPU_ABI_VERSION 2

/* error codes */
#define ERTGPU_OK 0
#define ERTGPU_EINVAL (-1)    /* bad argument / call order                       */
#define ERTGPU_ECUDA (-2)     /* CUDA runtime error, no device, wrong arch       */
#define ERTGPU_ENOMEM (-3)    /* host or device allocation failed                */
#define ERTGPU_ECAPACITY (-4) /* output array too small: *n_out holds the need   */
#define ERTGPU_ESIZE (-5)     /* nbytes not a multiple of BlockSize2 / too large */

#define ERTGPU_MAX_PROTOCOLS 8
#define ERTGPU_MAX_PREAMBLE 32
#define ERTGPU_MAX_PACKET_BYTES 92 /* (PacketSymbols+7)>>3 for idm/netidm */
#define ERTGPU_R900_DIGITS 42      /* r900.PayloadSymbols, r900/r900.go:30 */

/* integrity screens the GPU can run per candidate (the parsers re-check on the
 * host exactly as the reference does; the mask only lets the shim skip the
 * candidates every parser would reject) */
#define ERTGPU_CHECK_NONE 0
#define ERTGPU_CHECK_CRC16 1 /* Checksum(init, Bytes[from:to]) == residue: scm/scm.go:76, scmplus/scmplus.go:77 */
#define ERTGPU_CHECK_IDM 2   /* packet CRC Bytes[4:92] and serial CRC Bytes[9:13]+Bytes[88:90]: idm/idm.go:77-87, netidm/netidm.go:88-98 */
#define ERTGPU_CHECK_R900 3  /* base-6 digit pairs <= 31 and RS syndrome == 0: r900/r900.go:199-221 */

/* One registered parser's PacketConfig (protocol/decode.go:27-42) as the parser's
 * NewParser literal fills it (scm/scm.go:42-50 ...), plus the integrity screen. */
typedef struct {
    char name[16];                          /* PacketConfig.Protocol          */
    char preamble[ERTGPU_MAX_PREAMBLE + 1]; /* PacketConfig.Preamble, '0'/'1' */
    int32_t data_rate;                      /* PacketConfig.DataRate          */
    int32_t chip_length;                    /* PacketConfig.ChipLength (the -symbollength flag, main.go:77) */
    int32_t preamble_symbols;               /* PacketConfig.PreambleSymbols   */
    int32_t packet_symbols;                 /* PacketConfig.PacketSymbols     */
    uint32_t center_freq;                   /* PacketConfig.CenterFreq        */
    int32_t check_kind;                     /* ERTGPU_CHECK_*                 */
    uint16_t crc_init, crc_poly, crc_residue; /* crc.NewCRC arguments, crc/crc.go:16-24 */
    uint16_t reserved;
    int32_t crc_from, crc_to;               /* byte range for ERTGPU_CHECK_CRC16 */
} ertgpu_protocol;

/* Decoder.Cfg after Allocate (protocol/decode.go:131-141) */
typedef struct {
    int32_t data_rate;
    int32_t block_size, block_size2;
    int32_t chip_length, symbol_length;
    int32_t sample_rate;
    int32_t preamble_symbols, packet_symbols;
    int32_t preamble_length, packet_length;
    int32_t buffer_length;
    uint32_t center_freq;
    int32_t n_protocols;
    int32_t n_preambles;   /* distinct preambles, decode.go:121-124 */
    int32_t packet_bytes;  /* (packet_symbols+7)>>3, decode.go:154 */
} ertgpu_decoder_config;

/* One protocol.Data (protocol/parse.go:55-59) as produced by Search+Slice
 * (decode.go:255-375), tagged with the Decode call ("block") it belongs to. */
typedef struct {
    int64_t block;        /* index of the reference Decode call since stream start */
    int32_t idx;          /* Data.Idx in [0, BlockSize)                            */
    int32_t preamble_id;  /* distinct preamble, in registration order              */
    uint32_t check_mask;  /* bit i: i-th registered protocol passed its screen     */
    uint32_t flags;       /* ERTGPU_CAND_*                                         */
    uint8_t bytes[ERTGPU_MAX_PACKET_BYTES]; /* Data.Bytes (merged PacketSymbols, MSB first; a trailing partial byte holds its PK%8 bits in the low bits, like decode.go:363-366 on a zeroed pkt) */
    uint8_t r900_digits[ERTGPU_R900_DIGITS]; /* r900 quantized[] at the 42 payload positions, r900/r900.go:187-193 */
    uint8_t pad[2];
} ertgpu_candidate;

#define ERTGPU_CAND_HAS_R900 1u /* r900_digits[] is filled */

/* decode flags */
#define ERTGPU_DECODE_ONLY_VALID 1u /* return only candidates with check_mask != 0 */

/* parity taps (state of the most recent decode call) */
#define ERTGPU_TAP_SIGNAL 0    /* float32[BS+SL]   Decoder.Signal  after block b (decode.go:144,169)  */
#define ERTGPU_TAP_CSUM 1      /* float32[BS+SL+1] Decoder.csum    after block b (decode.go:147,232-236) */
#define ERTGPU_TAP_QUANTIZED 2 /* uint8[BUF]       Decoder.Quantized after block b (decode.go:145,166,243) */
#define ERTGPU_TAP_PACKED 3    /* uint8[(BS+PL+7)>>3] Decoder.packed after block b (decode.go:159,259-265) */
#define ERTGPU_TAP_R900_QUANTIZED 4 /* uint8[BUF]  r900 Parser.quantized after block b (r900/r900.go:82-150) */

typedef struct ertgpu_handle ertgpu_handle;

/* ---- lifecycle -------------------------------------------------------- */

/* protocol.NewDecoder (decode.go:65-71).  No CUDA work yet. */
int ertgpu_create(ertgpu_handle **out);

/* Decoder.RegisterProtocol (decode.go:100-128): merge the parser's config
 * (max of DataRate/ChipLength/PreambleSymbols/PacketSymbols, last CenterFreq)
 * and file it under its preamble.  Must precede ertgpu_allocate. */
int ertgpu_register_protocol(ertgpu_handle *h, const ertgpu_protocol *p);

/* Convenience: fill *p with the stock PacketConfig + screen of a parser by its
 * -msgtype name ("scm","scm+","idm","netidm","r900","r900bcd";
 * protocol/parse.go:42-51 NewParser).  Returns ERTGPU_EINVAL for unknown names
 * (the reference returns an error there, parse.go:49). */
int ertgpu_stock_protocol(const char *msgtype, int32_t chip_length, ertgpu_protocol *p);

/* Decoder.Allocate (decode.go:131-160): derive SymbolLength, BlockSize, ...,
 * select the CUDA device, allocate device buffers for up to
 * max_blocks_per_call blocks per decode call and max_candidates candidates.
 * max_blocks_per_call <= 0 or max_candidates <= 0 pick defaults. */
int ertgpu_allocate(ertgpu_handle *h, int32_t device, int64_t max_blocks_per_call,
                    int64_t max_candidates);

/* Decoder.Cfg (decode.go:46), valid after ertgpu_allocate. */
int ertgpu_get_config(const ertgpu_handle *h, ertgpu_decoder_config *cfg);

/* Forget the stream history (Signal tail, Quantized history, block counter):
 * the next decode behaves like the first Decode of a fresh Decoder. */
int ertgpu_reset(ertgpu_handle *h);

void ertgpu_destroy(ertgpu_handle *h);

const char *ertgpu_last_error(const ertgpu_handle *h);
int ertgpu_abi_version(void);

/* ---- decode ----------------------------------------------------------- */

/* N consecutive Decoder.Decode(input) calls (decode.go:163-197, called from
 * main.go:235) in one go: iq holds nbytes = N*BlockSize2 interleaved uint8 IQ
 * bytes in HOST memory.  Semantically identical to N sequential reference
 * Decodes: history (Signal tail, Quantized) is carried inside the handle
 * between calls.  Candidates (all preambles) are written in ascending
 * (block, preamble_id, idx) order.  Synchronous; copies are pipelined with
 * the kernels internally.  If more than cap candidates exist the call returns
 * ERTGPU_ECAPACITY with *n_out = the number needed and the stream state
 * already advanced (call ertgpu_fetch with a larger array to get them). */
int ertgpu_decode(ertgpu_handle *h, const uint8_t *iq, size_t nbytes, uint32_t flags,
                  ertgpu_candidate *out, size_t cap, size_t *n_out);

/* Same, with the IQ bytes already in DEVICE memory of the handle's device.
 * `stream` is a cudaStream_t (NULL = the handle's own stream).  Enqueues all
 * kernels on that stream and returns without synchronising; the results are
 * read with ertgpu_fetch.  d_iq must stay valid until then. */
int ertgpu_decode_device_async(ertgpu_handle *h, const void *d_iq, size_t nbytes, uint32_t flags,
                               void *stream);

/* Wait for the decode enqueued last and copy its candidates out (sorted as
 * above).  May be repeated (e.g. after ERTGPU_ECAPACITY). */
int ertgpu_fetch(ertgpu_handle *h, ertgpu_candidate *out, size_t cap, size_t *n_out);

/* Number of candidates of the last decode, and how many passed a screen. */
int ertgpu_last_counts(ertgpu_handle *h, int64_t *n_candidates, int64_t *n_valid);

/* Kernels launched by the last decode call (for bench.py's gpu_launches). */
int64_t ertgpu_last_launches(const ertgpu_handle *h);

/* Per-stage device timing of the decode pipeline (CUDA events recorded on the
 * launching stream around each stage; for bench.py's roofline).  After a decode
 * with timing enabled, ms4[] = {demod (magnitude+filter+quantize+pack),
 * search, slice+screens (+r900 replay), history carry} of the last pipeline. */
int ertgpu_set_stage_timing(ertgpu_handle *h, int32_t enable);
int ertgpu_last_stage_ms(ertgpu_handle *h, float *ms4);
/* Mean of the same four stage times over every pipeline since timing was enabled. */
int ertgpu_stage_ms_mean(ertgpu_handle *h, float *ms4, int64_t *n_pipelines);

/* Parity tap: reference buffer `which` as it would be after the Decode of
 * block `block` (absolute index; must lie inside the last decode call).
 * Writes up to cap bytes to dst (host), *n_out = bytes of the full tap. */
int ertgpu_tap(ertgpu_handle *h, int32_t which, int64_t block, void *dst, size_t cap,
               size_t *n_out);

/* Test hook: variant 0 forces the generic (any chip length) demod kernel,
 * -1 restores the automatic choice of a chip-length-specialised kernel. */
int ertgpu_set_demod_variant(ertgpu_handle *h, int32_t variant);

/* Pinned host memory for callers that can use it (cudaHostAlloc/cudaFreeHost).  ertgpu_decode
 * accepts any host memory: pinned input is copied to the device directly; ordinary (pageable)
 * input -- a Go slice, the block buffer of main.go:166 -- is staged through pinned buffers the
 * handle owns (the copy into them runs on a few host threads and overlaps the H2D transfer of
 * the previous chunk), so the caller never has to know. */
int ertgpu_host_alloc(void **out, size_t nbytes);
int ertgpu_host_free(void *p);

/* Restrict the CALLING thread to the CPUs next to `device` (the PCI device's local_cpulist in
 * sysfs), so that the pinned buffers it allocates afterwards and the staging copies it runs are
 * NUMA-local to the GPU.  Optional; *ncpus = CPUs in the new mask, *numa_node = the device's
 * node (-1 unknown).  Returns ERTGPU_EINVAL when the topology cannot be read (nothing changed). */
int ertgpu_bind_host_thread(int32_t device, int32_t *ncpus, int32_t *numa_node);

/* Names of the kernel instantiations the last decode launched (for bench.py's roofline label). */
const char *ertgpu_last_kernels(const ertgpu_handle *h);

/* ---- one stream over several GPUs (main.go:207-235 is where the
 * reference would hang such a driver) ------------------------------------------------- */

/* Shard r of a stream of total_blocks reference blocks cut into nshards contiguous block-aligned
 * shards: the shard REPORTS the candidates of blocks [first_block, last_block) and must be FED
 * from first_fed_block (halo: ceil(PacketLength/BlockSize) blocks of Quantized history,
 * decode.go:166, plus one block of Signal lead-in, decode.go:165). */
typedef struct {
    int64_t first_block, last_block, first_fed_block;
} ertgpu_shard;
int ertgpu_plan_shards(int64_t total_blocks, int32_t nshards, int32_t block_size, int32_t packet_length,
                       ertgpu_shard *out /* [nshards] */);

/* N consecutive Decoder.Decode calls on a FRESH stream, spread over nhandles handles (normally
 * one per GPU; several on one GPU also work): handle r decodes shard r (halo + owned blocks) of
 * the host buffer on its own device from its own host thread, candidates of halo blocks are
 * dropped, block numbers are made global and the lists are concatenated in (block, preamble,
 * idx) order -- the same list one handle returns for the whole buffer after ertgpu_reset.
 * All handles must be allocated with the same protocols; each is reset first.  There is no
 * collective: shards are independent (every candidate is decided inside one shard + halo). */
int ertgpu_decode_sharded(ertgpu_handle *const *handles, int32_t nhandles, const uint8_t *iq, size_t nbytes,
                          uint32_t flags, ertgpu_candidate *out, size_t cap, size_t *n_out);

/* ---- synthetic input (bench/test tooling, not part of the reference API) -- */

/* One injected packet: chips are OOK "high"/"low" flags, MSB first. */
typedef struct {
    int64_t start_sample;  /* first sample of chip 0 in the global stream */
    int32_t n_chips;       /* <= 1536                                      */
    int32_t chip_length;   /* samples per chip                             */
    int16_t amp_i, amp_q;  /* carrier added to I,Q on high chips           */
    uint8_t chips[192];
    int32_t pad;
} ertgpu_synth_packet;

/* Fill d_out (device) with nsamples IQ samples of the synthetic stream
 * starting at global sample index first_sample: counter-based noise keyed by
 * (seed, sample index) plus the given packets (host array, sorted by
 * start_sample, non-overlapping).  Bit-identical to synth_reference_fill() in
 * include/ertgpu_synth.h on the host. */
int ertgpu_synth_fill(int32_t device, void *d_out, int64_t first_sample, int64_t nsamples,
                      uint64_t seed, const ertgpu_synth_packet *packets, int64_t npackets,
                      void *stream);

#ifdef __cplusplus
}
#endif
#endif
