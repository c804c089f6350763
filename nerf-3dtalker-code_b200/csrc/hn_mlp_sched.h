// hn_mlp_sched.h — static schedules of the fused MLP kernels.
//
// Every fused kernel (forward chain, data-gradient chain) is an interpreter of three small tables that
// are generated ONCE on the host from the layer list below and kept in __constant__ memory:
//   PackOp : which slice of which fp32 state-dict weight goes into weight unit u (16 KiB operand image)
//   MmaOp  : for weight unit u, in stream order: A-operand block, accumulator, barriers to wait/commit
//   EpiOp  : for every accumulator chunk, in completion order: bias, activation, destination
// The weight producer, the MMA issuer and the epilogue warps all walk the same tables, so the three
// roles cannot drift apart.  fg_CD_predictor layer list: NetWorks/models.py:29-59, forward :62-87.
//
// RGB_layer_0 has no activation (models.py:79-80: x = RGB_layer_0(x); x = relu(RGB_layer_1(cat[x, appea]))), so it never
// runs as a layer here: hn_pack_weights multiplies it into RGB_layer_1's hidden block once per weight version
// (W_f = W_R1[:, :384] W_R0, 192 x 384; hn_fold_bias adds W_R1[:, :384] b_R0 to RGB_layer_1's effective bias), the chains go
// FeaExt_module_7 -> RGB_layer_1 directly ("optional algebraic fusion": -147 456 MAC per ray*sample
// in each of the three passes, 12 fewer saved operand blocks per tile), and hn_unfuse_r0r1 maps dL/dW_f back to the two
// weight gradients (two 192 x 384 x 384 products per step).
#pragma once
#include <stdint.h>

namespace hn {

// indices into hn_weights_t.w / hn_mlp_bwd_weights_t.dw
enum { W_L0 = 0, W_L5 = 5, W_L7 = 7, W_DENSITY = 8, W_R0 = 9, W_R1 = 10, W_R2 = 11 };

constexpr int kPeBlk = 6;          // A-operand block id of the positional-encoding block (0..5: activation buffer)
constexpr int kUnitBytes = 16384;

struct PackOp {                    // unit(r,c) = W[w_idx][(row0 + (T? c : r)) * ld + col0 + (T? r : c)]
    int8_t w_idx;
    int8_t transposed;             // data-gradient units hold W^T
    int8_t l5_hidden;              // add the runtime column of the hidden block of FeaExt_module_5
    int8_t pad;
    int16_t row0, col0;
    int16_t valid_r, valid_c;      // rows/cols of the unit backed by weights (rest is zero)
};

struct MmaOp {
    uint8_t a_blk;                 // A operand block (0..5 activation buffer, 6 = PE block)
    uint8_t n8;                    // MMA N / 8
    uint8_t tmem_col8;             // accumulator column / 8
    uint8_t q;                     // accumulator barrier index
    uint8_t first;                 // 1: overwrite accumulator (first K block of this chunk)
    uint8_t commit;                // 1: accumulator chunk complete after this unit -> acc_full[q]
    uint8_t wait_src;              // 0 none, 1..3 a_ready[c-1], 4 pe_ready, 5 in_ready (bwd: input image landed)
    uint8_t wait_empty;            // 1: wait acc_empty[q] before this unit
    uint16_t unit;                 // index of the first weight unit inside its packed stream
    uint8_t nkb;                   // consecutive 64-wide K blocks (weight units / A blocks) covered by this op: 1 or 2
    uint8_t pad;
};

enum EpiKind : uint8_t {
    EPI_HIDDEN = 0,                // y = relu(acc + bias)            -> activation buffer
    EPI_LINEAR = 1,                // y = acc + bias                  -> activation buffer
    EPI_FEAT = 2,                  // y = acc + bias                  -> global feat
    // data-gradient kernel
    EPI_GRAD_MASK = 3,             // y = acc * relu_mask             -> gradient buffer
    EPI_GRAD_LINEAR = 4,           // y = acc                         -> gradient buffer
    EPI_GRAD_DENSITY = 5,          // y = (acc + dsigma*w_density) * relu_mask
    EPI_GRAD_PE = 6,               // positional-encoding gradient    -> per-ray camera gradients
};

struct EpiOp {
    uint8_t q;
    uint8_t tmem_col8;
    uint8_t width32;               // columns / 32
    uint8_t kind;
    uint8_t dst_blk;               // first activation-buffer block written
    uint8_t ready_idx;             // a_ready barrier to arrive on, 255 = none
    uint8_t density;               // fwd: 1 accumulate density dot, 2 = also finish it (last chunk of FeaExt_module_7)
    uint8_t pad;
    uint16_t bias_off;             // fwd: offset in the per-item bias row ; bwd: unused
    uint16_t col0;                 // first logical output column of this chunk
    uint16_t save_blk;             // first block of the save slot (act / grads buffer), 0xFFFF = none
    uint16_t mask_word;            // word offset in the per-sample mask row, 0xFFFF = none
};

constexpr int kFwdUnits = 152;     // 150 weight units + 2 empty ones that keep every ring stage a PAIR of units
constexpr int kFwdStages = kFwdUnits / 2;
constexpr int kFwdEpis = 28;
constexpr int kBwdUnitsMax = 176;
constexpr int kBwdEpisMax = 36;

// ---- forward chain: the A operand (activations) lives in TENSOR MEMORY; the two 256-column halves of TMEM alternate
// between "accumulators of this layer" and "inputs of this layer" from layer to layer (and so from tile to tile: 11 layers).
// Inside the accumulator half (base B) of a 384-wide layer:
//   phase A  chunks 0 and 1 are ONE accumulator of 256 columns [B, B+256): every K block is a single tcgen05.mma of N = 256
//            whose B operand is a ring stage = the two weight units (chunk 0 rows, chunk 1 rows) of that K block;
//   phase B  chunk 2 reuses [B, B+128) once chunk 0's epilogue has drained it (barrier p_free); a stage holds two K blocks;
//   outputs  (128 outputs -> 64 columns of packed f16 pairs): chunk 0 -> B+192 (upper half of chunk 1's accumulator, after
//            the other epilogue group has loaded it), chunk 1 -> B+128 (in place), chunk 2 -> B (in place).
// The next layer reads those three slots as its K blocks and accumulates in the OTHER half, whose last readers (this layer's
// MMAs) precede it in issue order and whose last epilogue loads precede this layer's outputs: no accumulator barrier needed.
// All TMEM columns in the tables are for even tiles; odd tiles use column ^ 256.
constexpr uint16_t kSrcSmem = 0x8000;   // a_src flag: K block comes from shared memory (low bits: block index)
constexpr uint16_t kSrcNone = 0x7FFF;   // second unit of the stage absent (phase A stage, or an empty unit)
constexpr uint16_t kNoCol = 0xFFFF;

struct StageOp {                   // 16 bytes, one per ring stage (two consecutive weight units), read as one 128-bit word
    uint16_t a_src0;               // TMEM column of the A K block of the first MMA group, or kSrcSmem | block
    uint16_t a_src1;               // same for the second MMA group (phase B), kSrcNone = none
    uint16_t acc_col;              // TMEM column of the accumulator
    uint8_t n8;                    // MMA N / 8 (32: two chunks at once, 24: RGB_layer_1, 16: one chunk)
    uint8_t first;                 // 1: the first MMA of the stage overwrites the accumulator
    uint8_t commit;                // chunks completed by this stage: 0, 1 or 2 -> acc_full of chunk, chunk + 1
    uint8_t chunk;                 // first completing chunk (epilogue op index)
    uint8_t wait_src;              // before the stage: 0 none, 1..3 a_ready[c-1], 4 pe_ready
    uint8_t wait_p;                // 1: wait p_free (chunk 0's accumulator drained) before the stage
    uint32_t pad;
};
static_assert(sizeof(StageOp) == 16, "StageOp is read as one 128-bit word");

struct EpiOp2 {                   // 16 bytes: read with one 128-bit constant load
    uint16_t acc_col;              // TMEM column of the accumulator chunk
    uint16_t out_col;              // TMEM column of the packed output (64 columns per 128 outputs), kNoCol = none
    uint16_t bias_off;
    uint16_t col0;                 // first logical output column of this chunk
    uint16_t save_blk;             // first block of the save slot, 0xFFFF = none
    uint16_t mask_word;            // word offset in the per-sample mask row, 0xFFFF = none
    uint8_t width32;               // columns / 32
    uint8_t kind;
    uint8_t ready_idx;             // a_ready barrier to arrive on, 255 = none
    uint8_t flags;                 // bits 0-1: density (1 accumulate density dot, 2 = also finish it); bit 2: the output slot is
                                   // the accumulator of chunk n+1: wait for the other group's loads; bit 3: signal p_free after
                                   // this chunk's accumulator loads
};
static_assert(sizeof(EpiOp2) == 16, "EpiOp2 is read as one 128-bit word");

struct FwdTables { StageOp stage[kFwdStages]; EpiOp2 epi[kFwdEpis]; int n_stages; int pe_after_epi; int n_ready[3]; int n_epis;
                   int tile_flip;   // 1: odd tiles swap the TMEM halves (chains with an odd number of layers), 0: they do not
                   int pad; };

// ---- data-gradient chain without dL/dPE (the training step), same machinery as the forward chain: gradients live in tensor
// memory as the A operand, W^T units stream through the ring.  The first GEMM (RGB_layer_2^T) reads the four dL/dfeat
// K blocks from a two-block shared-memory ring (a_src = kSrcSmem | k, slot k & 1, every block awaited: wait_src 4).
constexpr int kBwdTUnits = 144;    // 2 * 72 stages (fits the forward tables' arrays: 72 <= 76 stages, 26 <= 28 chunks)
constexpr int kBwdTStages = kBwdTUnits / 2;
constexpr int kBwdTEpis = 26;

struct BwdTables { MmaOp mma[kBwdUnitsMax]; EpiOp epi[kBwdEpisMax]; int n_ops; int n_epis; int n_ready[3]; int n_empty[4]; };

struct HostSchedules {
    PackOp fwd_pack[kFwdUnits];
    FwdTables fwd;
    PackOp bwd_pack[kBwdUnitsMax];
    BwdTables bwd;                 // full data-gradient chain incl. dL/dPE (camera gradients)
    BwdTables bwd_nope;            // same without the dL/dPE chunks; `unit` still indexes the full stream
    int n_bwd_pack_units;          // weight units in the data-gradient stream (packed after the kFwdUnits forward units)
    PackOp bwdt_pack[kBwdTUnits];  // W^T units of the tensor-memory data-gradient chain (packed after the two streams above)
    FwdTables bwdt;
};

const HostSchedules& host_schedules();   // built on first use (hn_mlp_sched.cpp part of hn_mlp_pack.cu)

}  // namespace hn
