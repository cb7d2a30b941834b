// Shared definitions for libffn_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace ffn {

// ----------------------------------------------------------------------------------------
// Tile / shared-memory geometry of the fused render kernel
// ----------------------------------------------------------------------------------------
constexpr int kTileM = 128;                    // rows (samples) per tile == UMMA M == TMEM lanes
constexpr int kChunkK = 64;                    // K elements per 128-byte swizzled row
constexpr int kChunkBytesA = kTileM * 128;     // one K-chunk of an A tile: 16 KiB
constexpr int kActChunks = 4;                  // 256-wide hidden activation
constexpr int kEncChunk = 4;                   // chunk index of the encoding tile inside a slot
constexpr int kSlotChunks = 5;
constexpr int kSlotBytes = kSlotChunks * kChunkBytesA;   // 80 KiB
// every UMMA spans the CTA pair (cta_group::2): a CTA stages only ITS half of the rows of a weight K-chunk
constexpr int kWStageBytes = 128 * 128;        // half of a weight K-chunk (N=256 x K=64): 16 KiB per CTA
constexpr int kWStages = 4;
constexpr int kSmemW = 0;
constexpr int kSmemSlot0 = kWStages * kWStageBytes;             // 64 KiB
constexpr int kSmemMisc = kSmemSlot0 + 2 * kSlotBytes;          // 224 KiB
constexpr int kSmemMiscBytes = 3072;
constexpr int kSmemTotal = kSmemMisc + kSmemMiscBytes;          // 232448 = 227 KiB (the sm_100 maximum)

constexpr int kMaxMmaLayers = 12;
constexpr int kFoldedN = 128;                  // folded bottleneck . hidden_view layer of the inference program
// index of the weight-arena tensor map whose box is `rows` rows of 128 bytes (16 / 32 / 64 / 128)
__host__ __device__ __forceinline__ int wmap_index(uint32_t rows) {
  return rows >= 128 ? 3 : rows >= 64 ? 2 : rows >= 32 ? 1 : rows >= 16 ? 0 : 4;
}
constexpr int kMaxChunksPerLayer = 6;
// warps 0-3: weight producer / UMMA issuer / TMEM alloc / ones tile; 4-7, 8-11: epilogue warpgroup of slot 0 / 1.
// (A helper warpgroup per slot, lock-step weight sharing between the slots, single-CTA UMMAs and an A-operand-in-TMEM
// variant were measured in round 1 and removed: profiles/r01_perf_experiments.md.)
constexpr int kThreads = 384;

// epilogue kinds
enum : uint8_t {
  EPI_RELU_ACT = 0,     // act = relu(acc + b)        -> next A operand
  EPI_LINEAR_ACT = 1,   // act = acc + b              -> next A operand
  EPI_RELU_HEAD = 2,    // h = relu(acc + b); out[0 .. head_n) = head_w . h + head_b ; no A write
  EPI_ENC_PART2 = 3,    // no accumulator read: write the second half of a wide encoding into act chunks
  EPI_BWD_LINEAR = 4,   // backward: dz = acc                      -> next A operand (+ saved)
  EPI_BWD_MASK = 5,     // backward: dz = acc * relu'(h[mask_idx]) -> next A operand (+ saved)
};

// kernel passes
enum : int { PASS_INFER = 0, PASS_TRAIN_FWD = 1, PASS_BWD = 2 };

struct LayerDesc {
  uint32_t w_offset;                       // byte offset of this layer's packed weights
  uint16_t n;                              // UMMA N (256 or 128)
  uint8_t n_chunks;
  uint8_t epi;                             // EPI_*
  uint8_t accumulate;                      // 1: first MMA accumulates onto the existing TMEM tile
  uint8_t sigma_head;                      // 1: also emit out[3] = head_w[3] . h + head_b[3] from this layer's fp32 h;
                                           // 2 (folded inference program): columns [128,256) of that dot product are
                                           // read from TMEM AFTER the A operand has been handed over (the next layer is
                                           // 128 wide and leaves those accumulator columns intact)
  uint8_t write_view_enc;                  // 1: after this layer, overwrite the enc chunk with the view encoding
  uint8_t bias_row;                        // row of ConstParams::bias
  uint8_t src[kMaxChunksPerLayer];         // A chunk index (0..3 act, 4 enc) per K-chunk
  uint8_t ksteps[kMaxChunksPerLayer];      // UMMA K-steps (of 16) per K-chunk
  uint8_t head_n, has_bias, pad1, pad2;    // EPI_RELU_HEAD: outputs 0..head_n-1; has_bias: bias tile present
  uint32_t bias_off;                       // byte offset of the packed bias tile (N x 32 B, see kBiasTile*)
  uint32_t cbias_off;                      // byte offset of the COMPACT bias tile (N x 16 B: only the k < 8 core matrices;
                                           // the inference kernel keeps it outside the weight ring, see kSmemBiasBuf)
  int8_t save_idx;                         // training: slot of this layer's output in save_h / dz_out (-1 none)
  int8_t mask_idx;                         // training fwd: slot of the ReLU bitmask written; bwd: bitmask applied
  uint8_t pad3, pad4;
};

// Bias is folded into the accumulator by one extra K=16 UMMA per layer:
//   A = "ones" tile (every row [1,1,0,...,0]; 256 B in shared memory, all 16 row groups alias the same
//       two 8x16B core matrices through a stride-byte-offset of 0),
//   B = bias tile, K-major no-swizzle core matrices: B[n][0] = fp16(b_n), B[n][1] = fp16(b_n - B[n][0]).
// byte offset of element (n, k) inside a bias tile:
//   (n / 8) * kBiasTileSBO + (k / 8) * kBiasTileLBO + (n % 8) * 16 + (k % 8) * 2
constexpr int kBiasTileLBO = 128;   // between the two K-adjacent 8x8 core matrices
constexpr int kBiasTileSBO = 256;   // between N-adjacent core matrices (8-row groups)
constexpr int kSmemOnes = kSmemMisc + 2816;   // 256-byte ones tile at the end of the misc region
// Inference kernel: the bias tile of the current layer lives in its own 2 KB buffer instead of occupying a 16 KB stage
// of the weight ring for the 128 cycles of its UMMA (a ring window that contains a bias stage covers 1150 instead of
// 1540 cycles of tensor time against a ~1.3 k-cycle refill).  Only k = 0, 1 of a bias tile are non-zero and the "ones"
// A tile is zero for k >= 2, so the k >= 8 core matrices are dropped: leading-byte-offset 0 aliases them onto the
// k < 8 ones.  One tile per LAYER serves both slots.
constexpr int kSmemBiasBuf = kSmemMisc + 512;            // 2 KB: (N/2) rows x 16 B per CTA
constexpr int kSmemBarBiasFull = kSmemMisc + 256;        // mbarriers of that buffer
constexpr int kSmemBarBiasEmpty = kSmemMisc + 264;
constexpr int kCBiasTileSBO = 128;                       // between N-adjacent core matrices of the compact tile

// encodings
enum : int32_t { ENC_NERF = 0, ENC_FFMLP = 1, ENC_NONE = 2 };

// input modes
enum : int32_t { MODE_POINTS = 0, MODE_SAMPLES = 1, MODE_RAYS = 2, MODE_RAYS_T = 3 };

// Broadcast-read parameters: head weights (sigma / rgb / final layer) and frequency tables.  Lives in
// __constant__ memory; re-uploaded (device-to-device, stream ordered) whenever the
// active net or its weights change.
struct ConstParams {
  float head_w[4][256];
  float head_b[4];
  float freq_pos[16];
  float freq_view[16];
};

struct KernelArgs {
  const uint8_t* wpack;
  uint32_t wpack_lo_off;       // fp16x3 mode: byte offset of the residual ("lo") weight image inside the arena
  LayerDesc layers[kMaxMmaLayers];
  int32_t num_layers;
  int32_t enc_kind;
  int32_t f_pos, f_view, include_inputs, use_view;
  int32_t emb;                 // FFMLP embedding size E
  const float* ffm_b;          // FFMLP: device (3,E) b_values
  const float* ffm_a;          // FFMLP: device (E) a_values
  int32_t bf16;
  // inputs
  int32_t mode;
  const float* pos;            // POINTS (M,3) / SAMPLES (R,S,3)
  const float* dir;            // POINTS (M,3) / SAMPLES (R,S,3) / RAYS (R,3)
  const float* tvals;          // SAMPLES (R,S)
  const float* org;            // RAYS (R,3)
  const float* near_;          // RAYS (R)
  const float* far_;           // RAYS (R)
  const float* lin;            // RAYS (S)
  const float* jitter;         // RAYS (R,S) or null
  unsigned long long seed;
  long long ray_offset;
  int32_t stratified;
  long long M;                 // total rows
  int32_t S;                   // samples per ray (fused / rays modes)
  int32_t fused;               // 1: composite in-kernel (inference: any S; training forward: S power of two <= 128)
  float* ray_part;             // inference, rays that straddle 128-row tiles (128 % S != 0): [R][nseg_max][8] partials
  int32_t* ray_cnt;            //   ... and [R] arrival counters (zeroed before the launch)
  int32_t nseg_max;            //   tiles a ray can touch: (S + 126) / 128 + 1
  // outputs
  float* raw;                  // (M,4) when !fused
  float* t_out;                // (M) optional (RAYS)
  float* rgb;                  // (R,3)
  float* alpha;                // (R)
  float* depth;                // (R) or null
  int32_t* nan_flag;
  // debug
  int32_t dbg_layer;
  int32_t dbg_flags;           // bit 0: swap LBO/SBO roles of the no-swizzle descriptors (bring-up aid)
  float* dbg_out;              // (M,256)
  unsigned long long* stats;   // optional [32]: cycle counters (FFN_STATS=1), see ffn_debug_stats
  // training (PASS_TRAIN_FWD writes, PASS_BWD reads masks / writes dz)
  __nv_bfloat16* save_h;       // [n_save][M][256] layer outputs, post-activation, in the OPERAND dtype (fp16 | bf16)
  uint32_t* save_mask;         // [n_mask][M][8]   ReLU sign words (bit 31-j of word b <-> column 32b+j is <= 0)
  __half* save_enc;            // [2][M][64]       position / view encoding rows (our column order), operand dtype
  const float* d_raw;          // PASS_BWD: (M,4) gradient w.r.t. the raw network outputs [rgb | sigma]
  __nv_bfloat16* dz_out;       // PASS_BWD: [n_dz][M][256] gradients w.r.t. the pre-activations
  int32_t bwd_first_cols;      // PASS_BWD: width of the first dz tile (128 NeRF hidden_view, 256 FourierFeatureMLP)
  int32_t bwd_first_heads;     //           output heads feeding it (3 rgb | 4)
  int32_t bwd_first_mask;      //           sign-mask slot of that hidden layer
  int32_t bwd_first_save;      //           dz_out slot it is written to
  int32_t bwd_sigma_chunk;     //           1: d(sigma_raw) goes to column 0 of the encoding chunk
  // PASS_TRAIN_FWD, FourierFeatureMLP / un-encoded MLP: the first layer's input (the encoding, our column order)
  // is saved as two extra 256-column slots of save_h: slot x0_slot = encoding chunks 0..3 (features 0..127); slot
  // x0_slot + 1 = columns [0,192): chunks 5..7 (features 160..255), columns [192,256): chunk 4 (features 128..159, or
  // the raw inputs of the un-encoded MLP).  x0_slot < 0: nothing to save (NeRF: save_enc).
  int32_t x0_slot, x0_n1, x0_enc, x0_n2;
  int32_t num_tiles;
  int32_t dz_tma;              // PASS_BWD: 1 = dz_out is written by TMA stores of the bf16 A tile (dz_map)
  int32_t sh_tma;              // PASS_TRAIN_FWD: 1 = save_h is written by TMA stores of the A tile (sh_map)
  // the packed weight arena as rows of 128 bytes, one map per box height (16 / 32 / 64 / 128 rows = the half of a bias
  // tile or weight K-chunk a CTA stages; map 4: 8 rows, half of a 128-wide layer's compact bias tile): 2-SM TMA loads
  // that signal the leader CTA's barrier (ffn_pipeline.cuh)
  alignas(64) CUtensorMap wmap[5];
  alignas(64) CUtensorMap dz_map;   // [n_dz][M][256] bf16, boxes of 64 columns x 32 rows, SWIZZLE_128B
  alignas(64) CUtensorMap sh_map;   // [n_save][M][256] operand dtype, same boxes
};

}  // namespace ffn
