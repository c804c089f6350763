// hn_mlp_sched.h — static schedules of the fused MLP kernels.
//
// Every fused kernel (forward chain, data-gradient chain) is an interpreter of three small tables that
// are generated ONCE on the host from the layer list below and kept in __constant__ memory:
//   PackOp : which slice of which fp32 state-dict weight goes into weight unit u (16 KiB operand image)
//   MmaOp  : for weight unit u, in stream order: A-operand block, accumulator, barriers to wait/commit
//   EpiOp  : for every accumulator chunk, in completion order: bias, activation, destination
// The weight producer, the MMA issuer and the epilogue warps all walk the same tables, so the three
// roles cannot drift apart.  fg_CD_predictor layer list: NetWorks/models.py:29-59, forward :62-87.
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

constexpr int kFwdUnits = 168;
constexpr int kFwdEpis = 31;
constexpr int kBwdUnitsMax = 176;
constexpr int kBwdEpisMax = 36;

// ---- forward chain, second generation: the A operand (activations) lives in TENSOR MEMORY.
// TMEM (512 columns) is managed as eight 64-column slots.  A 128-wide accumulator chunk takes an aligned slot pair;
// its epilogue packs the 128 fp32 columns to 64 columns of f16 pairs (two 64-wide K blocks of the next GEMM) and
// stores them into a free slot, or in place over the first slot of its own accumulator.  The slot of every chunk is
// chosen on the host by a small allocator that knows the only two ordering facts the kernel provides:
//   (a) tcgen05.mma instructions execute in issue order (a later MMA may overwrite what an earlier one read);
//   (b) the MMA issuer waits, before chunk n, for the epilogue of chunk n-2 to have drained its accumulator; the
//       epilogue warps form two groups that take alternate chunks, so the stores of chunk n are ordered after the
//       accumulator loads of chunk n (own group barrier) and of every chunk <= n-2 (through (a) and acc_full of n);
//       a store into a slot drained by chunk n-1 (the other group) first waits for that group (EpiOp2::wait_prev).
// Chunks are computed N-outer (one chunk over all its K blocks, then the next), so the weight units stream in
// (layer, chunk, K block) order, one 16 KiB unit per ring stage.
constexpr uint16_t kSrcSmem = 0x8000;   // MmaOp2::a_src flag: K block comes from shared memory (low bits: block index)
constexpr uint16_t kNoCol = 0xFFFF;

struct MmaOp2 {
    uint16_t a_src;                // TMEM column of the 32-column A K block, or kSrcSmem | shared-memory block
    uint16_t acc_col;              // TMEM column of the accumulator chunk
    uint8_t n8;                    // MMA N / 8
    uint8_t first;                 // 1: first K block of the chunk (wait for the accumulator, overwrite it)
    uint8_t commit;                // 1: last K block of the chunk -> acc_full
    uint8_t wait_src;              // 0 none, 1..3 a_ready[c-1], 4 pe_ready
};

struct EpiOp2 {
    uint16_t acc_col;              // TMEM column of the accumulator chunk
    uint16_t out_col;              // TMEM column of the packed output (64 columns per 128 outputs), kNoCol = none
    uint8_t width32;               // columns / 32
    uint8_t kind;
    uint8_t ready_idx;             // a_ready barrier to arrive on, 255 = none
    uint8_t density;               // 1 accumulate density dot, 2 = also finish it
    uint8_t wait_prev;             // 1: the output slot was drained by chunk n-1: wait for the other group's loads
    uint8_t pad;
    uint16_t bias_off;
    uint16_t col0;                 // first logical output column of this chunk
    uint16_t save_blk;             // first block of the save slot, 0xFFFF = none
    uint16_t mask_word;            // word offset in the per-sample mask row, 0xFFFF = none
};

struct FwdTables { MmaOp2 mma[kFwdUnits]; EpiOp2 epi[kFwdEpis]; int n_ops; int pe_after_epi; int n_ready[3]; };
struct BwdTables { MmaOp mma[kBwdUnitsMax]; EpiOp epi[kBwdEpisMax]; int n_ops; int n_epis; int n_ready[3]; int n_empty[4]; };

struct HostSchedules {
    PackOp fwd_pack[kFwdUnits];
    FwdTables fwd;
    PackOp bwd_pack[kBwdUnitsMax];
    BwdTables bwd;                 // full data-gradient chain incl. dL/dPE (camera gradients)
    BwdTables bwd_nope;            // same without the dL/dPE chunks; `unit` still indexes the full stream
    int n_bwd_pack_units;          // weight units in the data-gradient stream (packed after the kFwdUnits forward units)
};

const HostSchedules& host_schedules();   // built on first use (hn_mlp_sched.cpp part of hn_mlp_pack.cu)

}  // namespace hn
