// capi.inl -- the extern "C" entry points of include/pg_b200.h over Engine<PG_BACKEND>.
// Included once by engine.cu (PG_BACKEND = pg::CudaBackend: the shipped library) and once by tests/emu (host backend,
// test infrastructure only).
struct pg_ctx { pg::Engine<PG_BACKEND> e; };

static bool misaligned(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) != 0; }
// every entry point first makes the ctx's device current (a process may hold contexts on several GPUs)
#define PG_NEED_CTX(ctx) do { if (!(ctx)) return PG_ERR_ARG; const_cast<pg_ctx*>(ctx)->e.be.activate(); } while (0)
#define PG_ALIGNED(ctx, p, dev) do { if ((dev) && (p) && misaligned(p)) return (ctx)->e.fail(PG_ERR_ARG, "device pointers must be 32-byte aligned"); } while (0)

// No C++ exception may cross the C ABI (the hosts are Rust, C and ctypes): host-side containers of the engine can throw
// std::bad_alloc / std::length_error (segment images, permutation tables, pg_srs_powers' table of powers ...).
template <class F>
static int pg_guarded(pg_ctx* ctx, F&& body) {
    try { return body(); }
    catch (const std::bad_alloc&) { return ctx->e.fail(PG_ERR_OOM, "host allocation failed"); }
    catch (const std::length_error&) { return ctx->e.fail(PG_ERR_OOM, "host container size limit exceeded"); }
    catch (const std::exception& ex) { return ctx->e.fail(PG_ERR_STATE, std::string("unexpected C++ exception: ") + ex.what()); }
    catch (...) { return ctx->e.fail(PG_ERR_STATE, "unexpected C++ exception"); }
}
#define PG_TRY(ctx, expr) return pg_guarded(const_cast<pg_ctx*>(ctx), [&]() -> int { return (expr); })

extern "C" {

int pg_abi_version(void) { return PG_B200_ABI_VERSION; }

const char* pg_strerror(int code) {
    switch (code) {
        case PG_OK: return "ok";
        case PG_ERR_NON_EXISTING_INVERSE: return "NonExistingInverse";
        case PG_ERR_CUDA: return "CUDA error";
        case PG_ERR_ARG: return "bad argument";
        case PG_ERR_OOM: return "out of device memory";
        case PG_ERR_MIXED_BITS: return "bounds of one batch must share one bit width";
        case PG_ERR_NO_DEVICE: return "no usable CUDA device (sm_100a required)";
        case PG_ERR_STATE: return "invalid state";
        default: return "unknown error";
    }
}
const char* pg_last_error(const pg_ctx* ctx) { return ctx ? ctx->e.err.c_str() : "null ctx"; }

int pg_ctx_create(const pg_cfg* cfg, pg_ctx** out) {
    if (!cfg || !out) return PG_ERR_ARG;
    *out = nullptr;
    pg_ctx* c = new (std::nothrow) pg_ctx();
    if (!c) return PG_ERR_OOM;
    int rc = pg_guarded(c, [&]() -> int { return c->e.create(*cfg); });
    if (rc != PG_OK) { fprintf(stderr, "pg_ctx_create: %s (%s)\n", pg_strerror(rc), c->e.err.c_str()); delete c; return rc; }
    *out = c;
    return PG_OK;
}
void pg_ctx_destroy(pg_ctx* ctx) { if (ctx) { ctx->e.be.activate(); try { ctx->e.destroy(); } catch (...) {} delete ctx; } }
int pg_composer_reset(pg_ctx* ctx) { PG_NEED_CTX(ctx); PG_TRY(ctx, ctx->e.reset()); }
int pg_sync(pg_ctx* ctx) { PG_NEED_CTX(ctx); PG_TRY(ctx, ctx->e.sync_checked()); }

int pg_add_input_batch(pg_ctx* ctx, uint64_t n, const pg_fr* values, int on_device, pg_col* out) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, values, on_device);
    PG_TRY(ctx, ctx->e.add_input_batch(n, values, on_device, out));
}
int pg_range_check_batch(pg_ctx* ctx, const pg_fr* mn, const pg_fr* mx, uint64_t n_bounds, int on_device, pg_col witness, pg_col* out, uint64_t* num_bits) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, mn, on_device); PG_ALIGNED(ctx, mx, on_device);
    PG_TRY(ctx, ctx->e.range_batch(true, mn, mx, n_bounds, on_device, witness, out, num_bits));
}
int pg_max_bound_batch(pg_ctx* ctx, const pg_fr* mx, uint64_t n_bounds, int on_device, pg_col witness, pg_col* out, uint64_t* num_bits) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, mx, on_device);
    PG_TRY(ctx, ctx->e.range_batch(false, nullptr, mx, n_bounds, on_device, witness, out, num_bits));
}
int pg_maybe_equal_batch(pg_ctx* ctx, pg_col a, pg_col b, pg_col* out) { PG_NEED_CTX(ctx); PG_TRY(ctx, ctx->e.maybe_equal_batch(a, b, out)); }
int pg_is_non_zero_batch(pg_ctx* ctx, pg_col var, const pg_fr* value_assigned, int on_device, uint64_t* n_err, uint64_t* first_err) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, value_assigned, on_device);
    PG_TRY(ctx, ctx->e.is_non_zero_batch(var, value_assigned, on_device, n_err, first_err));
}
int pg_is_non_zero_batch_flags(pg_ctx* ctx, pg_col var, const pg_fr* value_assigned, int on_device, uint8_t* err_flags, int layout, uint64_t* n_err) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, value_assigned, on_device);
    PG_TRY(ctx, ctx->e.is_non_zero_flags(var, value_assigned, on_device, err_flags, layout, n_err));
}
int pg_select_zero_batch(pg_ctx* ctx, pg_col x, pg_col select, pg_col* out) { PG_NEED_CTX(ctx); PG_TRY(ctx, ctx->e.select_batch(false, x, select, out)); }
int pg_select_one_batch(pg_ctx* ctx, pg_col y, pg_col selector, pg_col* out) { PG_NEED_CTX(ctx); PG_TRY(ctx, ctx->e.select_batch(true, y, selector, out)); }
int pg_constrain_to_constant_batch(pg_ctx* ctx, pg_col a, const pg_fr* constant, uint64_t n_const, const pg_fr* pi, uint64_t n_pi, int on_device) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, constant, on_device); PG_ALIGNED(ctx, pi, on_device);
    PG_TRY(ctx, ctx->e.constrain_batch(a, constant, n_const, pi, n_pi, on_device));
}

int pg_range_gate_batch(pg_ctx* ctx, pg_col witness, uint32_t num_bits) { PG_NEED_CTX(ctx); PG_TRY(ctx, ctx->e.range_gate_batch(witness, num_bits)); }

int pg_check(pg_ctx* ctx, uint64_t* n_unsat, uint64_t* first_bad_row) { PG_NEED_CTX(ctx); PG_TRY(ctx, ctx->e.check(n_unsat, first_bad_row)); }
int pg_check_rows(pg_ctx* ctx, uint64_t n, const pg_fr* w_val, const pg_fr* sel, const pg_fr* pi, int on_device, uint64_t* n_unsat, uint64_t* first_bad_row) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, w_val, on_device); PG_ALIGNED(ctx, sel, on_device); PG_ALIGNED(ctx, pi, on_device);
    PG_TRY(ctx, ctx->e.check_rows(n, w_val, sel, pi, on_device, n_unsat, first_bad_row));
}

int pg_check_rows_ex(pg_ctx* ctx, uint64_t n, const pg_fr* w_val, const pg_fr* sel, const pg_fr* pi, const pg_fr* q_arith, const pg_fr* q_range,
                     int on_device, uint64_t* n_unsat, uint64_t* first_bad_row) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, w_val, on_device); PG_ALIGNED(ctx, sel, on_device); PG_ALIGNED(ctx, pi, on_device);
    PG_ALIGNED(ctx, q_arith, on_device); PG_ALIGNED(ctx, q_range, on_device);
    PG_TRY(ctx, ctx->e.check_rows(n, w_val, sel, pi, on_device, n_unsat, first_bad_row, q_arith, q_range));
}
int pg_template_get(uint32_t gadget, uint32_t num_bits, const pg_fr* a, const pg_fr* b, uint64_t* n_rows, uint64_t* n_vars, int64_t* w_ref, pg_fr* sel, uint32_t* gate) {
    try { return pg::template_get(gadget, num_bits, a, b, n_rows, n_vars, w_ref, sel, gate); } catch (...) { return PG_ERR_STATE; }
}
int pg_op_shape(uint32_t gadget, uint32_t num_bits, uint64_t* rows, uint64_t* vars) { return pg::op_shape(gadget, num_bits, rows, vars) ? PG_OK : PG_ERR_ARG; }
int pg_shard_plan(const pg_op* ops, uint64_t n_ops, uint32_t world, int policy, pg_op_shard* out) {
    try { return pg::shard_plan(ops, n_ops, world, policy, out); } catch (...) { return PG_ERR_STATE; }
}
int pg_comm_unique_id(uint8_t* id) { return id && PG_BACKEND::comm_unique_id(id) ? PG_OK : PG_ERR_CUDA; }
int pg_comm_init(pg_ctx* ctx, const uint8_t* id, uint32_t rank, uint32_t world) { PG_NEED_CTX(ctx); PG_TRY(ctx, ctx->e.comm_init(id, rank, world)); }
int pg_comm_destroy(pg_ctx* ctx) { PG_NEED_CTX(ctx); ctx->e.be.sync(); ctx->e.be.comm_destroy(); return PG_OK; }
int pg_check_sharded(pg_ctx* ctx, const pg_op_shard* mine, uint64_t n_ops, uint64_t* n_unsat, uint64_t* first_bad_row, uint64_t* n_err) {
    PG_NEED_CTX(ctx); PG_TRY(ctx, ctx->e.check_sharded(mine, n_ops, n_unsat, first_bad_row, n_err));
}
int pg_gather_column(pg_ctx* ctx, pg_col col, pg_fr* dst, uint64_t capacity, int dst_on_device, uint64_t* counts, uint64_t* total) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, dst, dst_on_device);
    PG_TRY(ctx, ctx->e.gather_column(col, dst, capacity, dst_on_device, counts, total));
}
int pg_gather_variables(pg_ctx* ctx, uint64_t call, pg_fr* dst, uint64_t capacity, int dst_on_device, uint64_t* total) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, dst, dst_on_device);
    PG_TRY(ctx, ctx->e.gather_variables(call, dst, capacity, dst_on_device, total));
}
int pg_poke_variable(pg_ctx* ctx, uint64_t var, const pg_fr* value) { PG_NEED_CTX(ctx); PG_TRY(ctx, ctx->e.poke_variable(var, value)); }
int pg_counts(const pg_ctx* ctx, uint64_t* n_rows, uint64_t* n_vars) {
    PG_NEED_CTX(ctx);
    if (n_rows) *n_rows = ctx->e.n_rows;
    if (n_vars) *n_vars = ctx->e.n_vars;
    return PG_OK;
}
int pg_col_info(const pg_ctx* ctx, pg_col col, uint64_t* n, uint64_t* first_var, uint64_t* stride) {
    PG_NEED_CTX(ctx);
    const pg::Column* c = ctx->e.column(col);
    if (!c) return PG_ERR_ARG;
    const pg::Segment& s = ctx->e.segs[c->seg];
    if (n) *n = c->n;
    if (first_var) *first_var = s.base_var + c->inst_off * s.t.n_vars + c->local;
    if (stride) *stride = s.t.n_vars;
    return PG_OK;
}
int pg_col_read(pg_ctx* ctx, pg_col col, uint64_t i0, uint64_t cnt, pg_fr* dst, int dst_on_device) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, dst, dst_on_device == 1);
    if (dst_on_device < 0 || dst_on_device > 2) return ctx->e.fail(PG_ERR_ARG, "pg_col_read: dst_on_device must be 0, 1 or 2");
    PG_TRY(ctx, ctx->e.col_read(col, i0, cnt, dst, dst_on_device));
}
int pg_read_variables(pg_ctx* ctx, uint64_t var0, uint64_t cnt, pg_fr* dst, int dst_on_device) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, dst, dst_on_device);
    PG_TRY(ctx, ctx->e.read_variables(var0, cnt, dst, dst_on_device));
}
int pg_materialize_rows(pg_ctx* ctx, uint64_t row0, uint64_t cnt, uint64_t* w_idx, pg_fr* w_val, pg_fr* sel, pg_fr* pi, int dst_on_device) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, w_val, dst_on_device); PG_ALIGNED(ctx, sel, dst_on_device); PG_ALIGNED(ctx, pi, dst_on_device);
    PG_TRY(ctx, ctx->e.materialize(row0, cnt, w_idx, w_val, sel, pi, dst_on_device));
}

int pg_materialize_gate_selectors(pg_ctx* ctx, uint64_t row0, uint64_t cnt, pg_fr* q_arith, pg_fr* q_range, int dst_on_device) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, q_arith, dst_on_device); PG_ALIGNED(ctx, q_range, dst_on_device);
    PG_TRY(ctx, ctx->e.gate_selectors(row0, cnt, q_arith, q_range, dst_on_device));
}

int pg_permutation(pg_ctx* ctx, uint64_t row0, uint64_t cnt, uint64_t* sigma, int dst_on_device) {
    PG_NEED_CTX(ctx);
    PG_TRY(ctx, ctx->e.permutation(row0, cnt, sigma, dst_on_device));
}
int pg_export_composer(pg_ctx* ctx, const char* path, uint64_t chunk_rows, uint32_t flags) { PG_NEED_CTX(ctx); PG_TRY(ctx, ctx->e.export_composer(path, chunk_rows, flags)); }
int pg_fft(pg_ctx* ctx, uint32_t log_n, int inverse, const pg_fr* src, pg_fr* dst, int on_device) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, src, on_device); PG_ALIGNED(ctx, dst, on_device);
    PG_TRY(ctx, ctx->e.fft(log_n, inverse, src, dst, on_device));
}
int pg_wire_polynomials(pg_ctx* ctx, uint32_t log_n, pg_fr* dst, int dst_on_device) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, dst, dst_on_device);
    PG_TRY(ctx, ctx->e.wire_polynomials(log_n, dst, dst_on_device));
}
int pg_msm(pg_ctx* ctx, uint64_t n, const pg_g1_affine* points, const pg_fr* scalars, pg_g1_affine* out, int on_device) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, points, on_device); PG_ALIGNED(ctx, scalars, on_device);
    PG_TRY(ctx, ctx->e.msm(n, points, scalars, out, on_device));
}
int pg_srs_powers(pg_ctx* ctx, const pg_fr* beta, const pg_g1_affine* base, uint64_t n, pg_g1_affine* out, int out_on_device) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, out, out_on_device);
    PG_TRY(ctx, ctx->e.srs_powers(beta, base, n, out, out_on_device));
}
int pg_g1_fixed_base_mul(pg_ctx* ctx, uint64_t n, const pg_g1_affine* base, const pg_fr* scalars, pg_g1_affine* out, int on_device) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, scalars, on_device); PG_ALIGNED(ctx, out, on_device);
    PG_TRY(ctx, ctx->e.g1_fixed_base_mul(n, base, scalars, out, on_device));
}
int pg_commit_wire_polynomials(pg_ctx* ctx, uint32_t log_n, const pg_g1_affine* powers_of_g, uint64_t n_powers, int powers_on_device, pg_g1_affine* out4) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, powers_of_g, powers_on_device);
    PG_TRY(ctx, ctx->e.commit_wire_polynomials(log_n, powers_of_g, n_powers, powers_on_device, out4));
}
int pg_srs_lagrange(pg_ctx* ctx, const pg_fr* beta, const pg_g1_affine* base, uint32_t log_n, pg_g1_affine* out, int out_on_device) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, out, out_on_device);
    PG_TRY(ctx, ctx->e.srs_lagrange(beta, base, log_n, out, out_on_device));
}
int pg_commit_wire_evaluations(pg_ctx* ctx, uint32_t log_n, const pg_g1_affine* lagrange, uint64_t n_points, int points_on_device, pg_g1_affine* out4) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, lagrange, points_on_device);
    PG_TRY(ctx, ctx->e.commit_wire_evaluations(log_n, lagrange, n_points, points_on_device, out4));
}
int pg_g1_op(pg_ctx* ctx, int op, uint64_t n, const pg_g1_affine* a, const pg_g1_affine* b, pg_g1_affine* out) { PG_NEED_CTX(ctx); PG_TRY(ctx, ctx->e.g1_op(op, n, a, b, out)); }
int pg_fr_to_bytes(pg_ctx* ctx, uint64_t n, const pg_fr* src, uint8_t* dst, int on_device) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, src, on_device); PG_ALIGNED(ctx, dst, on_device);
    PG_TRY(ctx, ctx->e.convert(true, n, src, reinterpret_cast<pg_fr*>(dst), on_device, nullptr, nullptr));
}
int pg_fr_from_bytes(pg_ctx* ctx, uint64_t n, const uint8_t* src, pg_fr* dst, int on_device, uint64_t* n_invalid, uint64_t* first_invalid) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, src, on_device); PG_ALIGNED(ctx, dst, on_device);
    PG_TRY(ctx, ctx->e.convert(false, n, reinterpret_cast<const pg_fr*>(src), dst, on_device, n_invalid, first_invalid));
}

int pg_synth(pg_ctx* ctx, uint64_t seed, uint64_t stream, uint64_t n, int kind, uint32_t bits, pg_fr* dst_device) {
    PG_NEED_CTX(ctx); PG_ALIGNED(ctx, dst_device, 1);
    PG_TRY(ctx, ctx->e.synth(seed, stream, n, kind, bits, dst_device));
}
int pg_get_timing(pg_ctx* ctx, pg_timing* out, int reset) {
    PG_NEED_CTX(ctx);
    if (!out) return PG_ERR_ARG;
    return ctx->e.be.timing(out, reset != 0) ? PG_OK : ctx->e.fail(PG_ERR_CUDA, "timing");
}
int pg_get_check_stats(pg_ctx* ctx, pg_check_stats* out, int reset) {
    PG_NEED_CTX(ctx);
    if (!out) return PG_ERR_ARG;
    *out = ctx->e.be.ck;
    if (reset) ctx->e.be.ck = pg_check_stats{};
    return PG_OK;
}
int pg_measure_imad_peak(pg_ctx* ctx, double* wide_mac_per_s, double* imad_per_s) {
    PG_NEED_CTX(ctx);
    double w = 0, l = 0;
    if (!ctx->e.be.imad_peak(&w, &l)) return ctx->e.fail(PG_ERR_CUDA, "imad peak");
    if (wide_mac_per_s) *wide_mac_per_s = w;
    if (imad_per_s) *imad_per_s = l;
    return PG_OK;
}
int pg_microbench(pg_ctx* ctx, int mode, double* ops_per_s) {
    PG_NEED_CTX(ctx);
    double v = 0;
    if (!ctx->e.be.ubench(mode, &v)) return ctx->e.fail(PG_ERR_CUDA, "microbench");
    if (ops_per_s) *ops_per_s = v;
    return PG_OK;
}
int pg_fr_op(pg_ctx* ctx, int op, uint64_t n, const pg_fr* a, const pg_fr* b, pg_fr* out) { PG_NEED_CTX(ctx); PG_TRY(ctx, ctx->e.fr_op(op, n, a, b, out)); }

}  // extern "C"
