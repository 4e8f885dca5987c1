// quotient.cu — device side of halo2's `evaluate_h` (UPSTREAM plonk/evaluation.rs: Evaluator::evaluate_h / GraphEvaluator),
// the quotient-numerator stage of create_proof (SURVEY.md §3.1 stage 6, entered from /root/reference/src/pfsys/mod.rs:456).
//
// halo2 compiles every gate, permutation and lookup constraint into a straight-line program of field operations over
// "value sources" (column cosets at a rotation, constants, challenges, earlier intermediates) and runs it once per row of
// the extended domain.  This kernel is that interpreter: one thread per extended row, the program broadcast from global
// memory (uniform control flow, no divergence), intermediates in a small per-thread slot file.  Row rotations are cyclic
// index offsets (rotation * 2^(ext_k - k)) resolved on the host.  The Rust side lowers its GraphEvaluator calculations to
// QInstr (include/ezkl_b200.h: b200_instr); l0 / l_last / l_active_row, the identity coset X and previous partial sums are
// ordinary columns, y / beta / gamma / theta are constants.  HBM traffic per row: 32 B per distinct (column, rotation)
// load + 32 B store; arithmetic is bound by the same multiply ceiling as every other kernel here.
#include <vector>
#include "quotient.cuh"

namespace b200 {

__global__ void __launch_bounds__(128) k_quotient_eval(const Fr* const* __restrict__ cols, uint32_t mask, const QLoad* __restrict__ loads,
                                                        const Fr* __restrict__ consts, const QInstr* __restrict__ prog, uint32_t n_instr, Fr* __restrict__ out) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx > mask) return;
    Fr slots[Q_MAX_SLOTS];
    uint32_t last = 0;
#pragma unroll 1
    for (uint32_t pc = 0; pc < n_instr; ++pc) {
        const QInstr in = prog[pc];
        const uint32_t op = in.op_dst & 0xff, dst = (in.op_dst >> 8) & (Q_MAX_SLOTS - 1);
        Fr x, y;
        {
            const uint32_t k = in.a >> 30, i = in.a & 0x3fffffffu;
            if (k == QSRC_SLOT) x = slots[i & (Q_MAX_SLOTS - 1)];
            else if (k == QSRC_CONST) x = fp_load(consts + i);
            else { const QLoad l = loads[i]; x = fp_load(cols[l.column] + ((idx + l.offset) & mask)); }
        }
        if (op <= QOP_MUL) {
            const uint32_t k = in.b >> 30, i = in.b & 0x3fffffffu;
            if (k == QSRC_SLOT) y = slots[i & (Q_MAX_SLOTS - 1)];
            else if (k == QSRC_CONST) y = fp_load(consts + i);
            else { const QLoad l = loads[i]; y = fp_load(cols[l.column] + ((idx + l.offset) & mask)); }
        }
        Fr r;
        switch (op) {
            case QOP_ADD: r = x + y; break;
            case QOP_SUB: r = x - y; break;
            case QOP_MUL: r = x * y; break;
            case QOP_NEG: r = fp_neg(x); break;
            case QOP_DOUBLE: r = fp_dbl(x); break;
            case QOP_SQUARE: r = fp_sqr(x); break;
            default: r = x; break;
        }
        slots[dst] = r;
        last = dst;
    }
    fp_store(out + idx, n_instr ? slots[last] : fp_zero<FrTag>());
}

int quotient_eval_run(const Fr* const* h_col_ptrs, size_t n_cols, uint32_t ext_k, const QLoad* h_loads, size_t n_loads, const Fr* h_consts, size_t n_consts,
                      const QInstr* h_prog, size_t n_instr, Fr* d_out, QuotientWorkspace& ws, cudaStream_t st) {
    B200_CHECK(ext_k >= 1 && ext_k <= 28, -1, "quotient_eval: ext_k %u out of range", ext_k);
    B200_CHECK(n_instr < (1u << 24) && n_loads < (1u << 30) && n_consts < (1u << 30), -1, "quotient_eval: program too large");
    const uint32_t N = 1u << ext_k;
    // validate the program on the host so the kernel can index without checks
    for (size_t i = 0; i < n_loads; ++i) B200_CHECK(h_loads[i].column < n_cols && h_loads[i].offset < N, -1, "quotient_eval: load %zu out of range", i);
    for (size_t pc = 0; pc < n_instr; ++pc) {
        const uint32_t op = h_prog[pc].op_dst & 0xff, dst = h_prog[pc].op_dst >> 8;
        B200_CHECK(op <= QOP_MOV && dst < (uint32_t)Q_MAX_SLOTS, -1, "quotient_eval: instruction %zu: bad op %u / slot %u", pc, op, dst);
        const uint32_t srcs[2] = {h_prog[pc].a, h_prog[pc].b};
        for (int s = 0; s < (op <= QOP_MUL ? 2 : 1); ++s) {
            const uint32_t k = srcs[s] >> 30, i = srcs[s] & 0x3fffffffu;
            B200_CHECK(k <= QSRC_LOAD && ((k == QSRC_SLOT && i < (uint32_t)Q_MAX_SLOTS) || (k == QSRC_CONST && i < n_consts) || (k == QSRC_LOAD && i < n_loads)), -1,
                       "quotient_eval: instruction %zu: bad operand", pc);
        }
    }
    // one staging blob: [col ptrs | loads | consts | prog]
    const size_t o_cols = 0, o_loads = (o_cols + sizeof(void*) * n_cols + 31) & ~(size_t)31, o_consts = (o_loads + sizeof(QLoad) * n_loads + 31) & ~(size_t)31,
                 o_prog = (o_consts + sizeof(Fr) * n_consts + 31) & ~(size_t)31, total = o_prog + sizeof(QInstr) * n_instr + 32;
    std::vector<uint8_t> blob(total, 0);
    if (n_cols) memcpy(blob.data() + o_cols, h_col_ptrs, sizeof(void*) * n_cols);
    if (n_loads) memcpy(blob.data() + o_loads, h_loads, sizeof(QLoad) * n_loads);
    if (n_consts) memcpy(blob.data() + o_consts, h_consts, sizeof(Fr) * n_consts);
    if (n_instr) memcpy(blob.data() + o_prog, h_prog, sizeof(QInstr) * n_instr);
    uint8_t* d = reinterpret_cast<uint8_t*>(ws.ring.push(blob.data(), total, st));
    if (!d) {
        if (ws.prog.ensure(total)) return -2;
        B200_CUDA(cudaMemcpyAsync(ws.prog.p, blob.data(), total, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaStreamSynchronize(st));      // blob is a stack temporary
        d = ws.prog.as<uint8_t>();
    }
    k_quotient_eval<<<div_up(N, 128), 128, 0, st>>>(reinterpret_cast<const Fr* const*>(d + o_cols), N - 1, reinterpret_cast<const QLoad*>(d + o_loads),
                                                     reinterpret_cast<const Fr*>(d + o_consts), reinterpret_cast<const QInstr*>(d + o_prog), (uint32_t)n_instr, d_out);
    B200_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200
