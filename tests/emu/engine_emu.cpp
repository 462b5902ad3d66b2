// tests/emu/engine_emu.cpp -- TEST INFRASTRUCTURE ONLY: the engine's host logic (engine.hpp, templates.hpp) and the
// per-instance kernel bodies (bodies.cuh) compiled by g++ over a host backend that runs the bodies in a loop.
//
// Purpose: check template construction, Variable/row numbering, operand resolution and witness arithmetic against the
// oracle on a machine without a GPU.  What it does NOT cover: the __global__ wrappers, the block-wide batch inversion
// (replaced here by one Fermat inversion per element) and the PTX carry chains (covered by fr_emu.cpp) -- those are
// checked on the B200 by the `-m gpu` tests.  This library is built into tests/emu/_build and loaded only by
// tests/test_emu_*.py; the package plonk_gadgets_b200 never loads it and has no CPU fallback.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include "../../plonk_gadgets_b200/csrc/engine.hpp"

namespace pg {

Fr h_pow2[256];

struct HostPool {
    const uint32_t* p;
    Fr operator()(uint32_t idx) const { return pool_load(p, idx); }
};

class HostBackend {
public:
    const char* error() const { return ""; }
    bool no_device() const { return false; }
    void activate() {}
    bool init(const pg_cfg&) { return true; }
    void shutdown() {}
    void* alloc(size_t bytes) { void* p = nullptr; if (posix_memalign(&p, 64, bytes ? bytes : 64)) return nullptr; memset(p, 0xA5, bytes); return p; }
    void release(void* p) { free(p); }
    bool h2d(void* d, const void* s, size_t n) { memcpy(d, s, n); return true; }
    bool h2d_chunked(void* d, const void* s, uint64_t n, size_t elem, uint64_t, unsigned long long* counters) {
        memcpy(d, s, n * elem);
        if (counters) { const ValidateBody::Args va{reinterpret_cast<const uint4*>(d), 0, n, counters}; for (uint64_t i = 0; i < n; i++) ValidateBody::run(va, i); }
        return true;
    }
    bool d2h(void* d, const void* s, size_t n) { memcpy(d, s, n); return true; }
    bool d2d(void* d, const void* s, size_t n) { memmove(d, s, n); return true; }
    bool d2h_async(void* d, const void* s, size_t n) { memcpy(d, s, n); return true; }
    bool sync() { return true; }
    bool upload_pow2(const Fr*) { return true; }
    bool timing(pg_timing* out, bool) { memset(out, 0, sizeof(*out)); return true; }
    bool imad_peak(double*, double*) { return false; }
    bool ubench(int, double*) { return false; }

    template <class Body>
    bool run_simple(const typename Body::Args& a, uint64_t n, int) {
        for (uint64_t i = 0; i < n; i++) Body::run(a, i);
        return true;
    }
    // the chunk-by-chunk launch of the CUDA backend (operand still arriving), here in three uneven chunks through the same i0 / n
    template <class Body>
    bool run_simple_chunked(const typename Body::Args& a_in, uint64_t n, int, const void*) {
        const uint64_t cut[4] = {0, n / 3, n / 3 + n / 2, n};     // n/3 + n/2 <= n
        for (int k = 0; k < 3; k++) {
            typename Body::Args a = a_in;
            a.i0 = cut[k]; a.n = cut[k + 1] - cut[k];
            for (uint64_t i = 0; i < a.n; i++) Body::run(a, i);
        }
        return true;
    }
    // the range gadgets' three kernels; in three uneven chunks (as behind a chunked input copy on the GPU) when n is large enough
    void drop_result_chunks() {}
    bool result_chunks_of(const void*) const { return false; }
    bool col_read_chunked(const ColReadBody::Args&, void*, const void*) { return false; }
    template <class Pre, class Post>
    bool run_range_pipeline(const typename Pre::Args& a_in, const BatchInvArgs& inv_in, uint64_t n, const void* operand_table, const void*, bool chunk_results) {
        if (n < 6 || !chunk_results) return run_simple_chunked<Pre>(a_in, n, 0, operand_table) && run_batch_inv(inv_in, 0) && run_simple<Post>(a_in, n, 0);
        const uint64_t cut[4] = {0, n / 3, n / 3 + n / 2, n};
        for (int k = 0; k < 3; k++) {
            typename Pre::Args a = a_in;
            a.i0 = cut[k]; a.n = cut[k + 1] - cut[k];
            BatchInvArgs inv = inv_in;
            inv.fr = inv_in.fr + 2 * cut[k]; inv.n = a.n;
            for (uint64_t i = 0; i < a.n; i++) Pre::run(a, i);
            if (!run_batch_inv(inv, 0)) return false;
            for (uint64_t i = 0; i < a.n; i++) Post::run(a, i);
        }
        return true;
    }
    bool run_batch_inv(const BatchInvArgs& a, int) {   // one Fermat inversion per element (the block-wide trick is GPU-only)
        for (uint32_t j = 0; j < a.n_pairs; j++)
            for (uint64_t i = 0; i < a.n; i++) {
                const Fr x = tab_load_fr(a.fr, a.stride, a.in_slot[j], i);
                tab_store_fr(a.fr, a.stride, a.out_slot[j], i, fr_is_zero(x) ? fr_zero() : fr_inv_fermat(x));
            }
        return true;
    }
    template <class Hook>
    bool run_batch_inv_fused(const BatchInvArgs& a, const typename Hook::Args& h, int) {   // single pair: element i = instance i
        for (uint64_t i = 0; i < a.n; i++) {
            const Fr x = Hook::pre(h, i);
            const Fr z = fr_is_zero(x) ? fr_zero() : fr_inv_fermat(x);
            tab_store_fr(a.fr, a.stride, a.out_slot[0], i, z);
            Hook::post(h, i, x, z);
        }
        return true;
    }
    // a world of one rank (the collectives are GPU-side: comm.cuh)
    const char* comm_error() const { return "the host backend has no communicator"; }
    int comm_rank() const { return 0; }
    int comm_world() const { return 1; }
    static bool comm_unique_id(uint8_t*) { return false; }
    bool comm_init(const uint8_t*, int, int) { return false; }
    void comm_destroy() {}
    bool comm_verdict(const unsigned long long* c, unsigned long long n_err, unsigned long long out[4], const FusedSpan* map, uint32_t n_map) {
        unsigned long long unsat = c[CNT_UNSAT], first = c[CNT_FIRST_BAD];
        if (n_map) {
            unsat += c[CNT_FUSED_UNSAT];
            unsigned long long ff = c[CNT_FUSED_FIRST];
            if (ff != ~0ull) {
                for (uint32_t k = 0; k < n_map; k++) if (ff >= map[k].local_base && ff < map[k].local_end) { ff = map[k].global_base + (ff - map[k].local_base); break; }
                if (ff < first) first = ff;
            }
        }
        out[0] = unsat; out[1] = n_err; out[2] = first; out[3] = c[CNT_BAD_INPUT];
        return true;
    }
    bool comm_counts(unsigned long long mine, unsigned long long* counts) { counts[0] = mine; return true; }
    bool comm_gather(const void* send, void* recv, const unsigned long long* counts) { memcpy(recv, send, (size_t)counts[0] * 32); return true; }
    pg_check_stats ck{};
    void count_check(int kind, uint64_t rows) { ck.launches[kind]++; ck.rows[kind] += rows; }
    bool run_check(const CheckArgs& a, const SparseProg& prog) {
        HostPool pool = {a.pool};
        const QRegs q = q_regs_default();
        count_check(a.n_inst < 48 && a.n_rows > 1 ? PG_CK_ROWPAR : !a.mode ? PG_CK_INSTANCE_GENERIC : prog.ops ? PG_CK_PROGRAM : PG_CK_INSTANCE_TERMS, a.n_inst * a.n_rows);
        if (a.n_inst < 48 && a.n_rows > 1) {            // the row-parallel mapping (thread = one row of one instance), as for small segments on the GPU
            for (uint64_t t = 0; t < a.n_inst * a.n_rows; t++) {
                unsigned long long fb = ~0ull;
                const uint32_t bad = a.mode ? CheckBody::run_one<1>(a, pool, q, t, fb) : CheckBody::run_one<0>(a, pool, q, t, fb);
                if (bad) { a.counters[CNT_UNSAT] += bad; if (fb < a.counters[CNT_FIRST_BAD]) a.counters[CNT_FIRST_BAD] = fb; }
            }
            return true;
        }
        for (uint64_t i = 0; i < a.n_inst; i++) {
            unsigned long long fb = ~0ull;
            const uint32_t bad = a.mode ? (prog.ops ? SparseProgBody::run<4>(a, prog, pool, QDefault(), i, fb) : CheckBody::run<1>(a, pool, q, i, fb)) : CheckBody::run<0>(a, pool, q, i, fb);
            if (bad) { a.counters[CNT_UNSAT] += bad; if (fb < a.counters[CNT_FIRST_BAD]) a.counters[CNT_FIRST_BAD] = fb; }
        }
        return true;
    }
    bool run_check_gates(const CheckArgs& a) {
        count_check(PG_CK_GATES, a.n_inst * a.n_rows);
        HostPool pool = {a.pool};
        const QRegs q = q_regs_default();
        if (a.n_inst >= 48) {                            // one thread per instance walking the rows, as for large segments on the GPU
            for (uint64_t i = 0; i < a.n_inst; i++) {
                unsigned long long fb = ~0ull;
                const uint32_t bad = GateRowsCheckBody::run(a, pool, q, i, fb);
                if (bad) { a.counters[CNT_UNSAT] += bad; if (fb < a.counters[CNT_FIRST_BAD]) a.counters[CNT_FIRST_BAD] = fb; }
            }
            return true;
        }
        for (uint64_t t = 0; t < a.n_inst * a.n_rows; t++) {
            unsigned long long fb = ~0ull;
            const uint32_t bad = GateRowsCheckBody::run_one(a, pool, q, t, fb);
            if (bad) { a.counters[CNT_UNSAT] += bad; if (fb < a.counters[CNT_FIRST_BAD]) a.counters[CNT_FIRST_BAD] = fb; }
        }
        return true;
    }
    bool run_mat_tiled(const MatTileArgs& a) {      // same outputs as the tiled CUDA kernel, element by element
        const DevSeg& s = a.seg;
        for (uint64_t ii = 0; ii < a.n_inst; ii++)
            for (uint32_t r = 0; r < s.n_rows; r++) {
                const uint64_t o = a.out_off + ii * s.n_rows + r;
                for (int w = 0; w < 4; w++) if (a.w_val) aos_store(a.w_val, (uint64_t)w * a.stride + o, loc_load(s.tab, s.rows[r].loc[w], a.inst0 + ii));
                for (int k = 0; k < 5; k++) if (a.sel) aos_store(a.sel, (uint64_t)k * a.stride + o, pool_load(s.pool, s.rows[r].sel[k]));
            }
        return true;
    }
    bool sort_pairs(const uint32_t* keys, const uint32_t* vals, uint32_t* keys_out, uint32_t* vals_out, uint64_t n, uint32_t) {
        std::vector<uint64_t> idx(n);
        for (uint64_t i = 0; i < n; i++) idx[i] = i;
        std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return keys[a] < keys[b]; });
        for (uint64_t i = 0; i < n; i++) { keys_out[i] = keys[idx[i]]; vals_out[i] = vals[idx[i]]; }
        return true;
    }
    bool run_msm_buckets(const MsmBucketBody::Args& a) { for (uint64_t b = 0; b < a.n; b++) MsmBucketBody::run(a, b); return true; }
    bool exclusive_sum(const uint32_t* in, uint32_t* out, uint64_t n) { uint32_t acc = 0; for (uint64_t i = 0; i < n; i++) { const uint32_t v = in[i]; out[i] = acc; acc += v; } return true; }
    bool run_ntt_pass(const NttPassArgs& a, uint64_t n_blocks) { ntt_pass_host(a, n_blocks); return true; }
    bool run_check_rows(const CheckRowsBody::Args& a) {
        for (uint64_t i = 0; i < a.n; i++)
            if (CheckRowsBody::run(a, i)) { a.counters[CNT_UNSAT]++; if (i < a.counters[CNT_FIRST_BAD]) a.counters[CNT_FIRST_BAD] = i; }
        return true;
    }
};

}  // namespace pg

#define PG_BACKEND pg::HostBackend
#include "../../plonk_gadgets_b200/csrc/capi.inl"
