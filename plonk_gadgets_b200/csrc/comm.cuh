// comm.cuh -- the two collectives of the sharded path, NCCL over NVLink 5 / NVSwitch, on the engine's stream:
//   * all-reduce of the verdict (sum of unsatisfied-row and error counts, min of the first bad row);
//   * all-gather of per-instance results / witness shards whose lengths differ between ranks.
// No collective sits inside a compute kernel: instances are independent, ranks only meet here.
//
// NCCL is bound at run time (dlopen "libnccl.so.2": the copy a host process has loaded already -- PyTorch ships one -- or the
// system's), so the library itself has no link-time dependency on it and single-GPU hosts never load it.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

namespace pg {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;

    // returns nullptr on success, else what failed
    const char* load() {
        if (handle) return nullptr;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (handle) break; }
        if (!handle) return "libnccl.so.2 not found (dlopen)";
#define PG_NCCL_SYM(field, name) do { *(void**)(&field) = dlsym(handle, name); if (!field) return "NCCL symbol missing: " name; } while (0)
        PG_NCCL_SYM(GetUniqueId, "ncclGetUniqueId"); PG_NCCL_SYM(CommInitRank, "ncclCommInitRank"); PG_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        PG_NCCL_SYM(AllReduce, "ncclAllReduce"); PG_NCCL_SYM(AllGather, "ncclAllGather"); PG_NCCL_SYM(Broadcast, "ncclBroadcast");
        PG_NCCL_SYM(GroupStart, "ncclGroupStart"); PG_NCCL_SYM(GroupEnd, "ncclGroupEnd"); PG_NCCL_SYM(GetErrorString, "ncclGetErrorString");
        PG_NCCL_SYM(GetVersion, "ncclGetVersion");
#undef PG_NCCL_SYM
        return nullptr;
    }
};
inline NcclApi& nccl_api() { static NcclApi api; return api; }

static_assert(sizeof(ncclUniqueId) == PG_COMM_ID_BYTES, "PG_COMM_ID_BYTES must be sizeof(ncclUniqueId)");

// pack the verdict words for the all-reduce: v[0] = n_unsat, v[1] = n_err (sum); v[2] = first bad row (min); v[3] = unreduced inputs.
// Rows that were evaluated inside witness generation (PG_F_FUSED_CHECK) recorded their verdict in CNT_FUSED_* with the local row
// numbering of the time; `map` lists those segments so that the first bad row is renumbered like the launched checks are.
__global__ void k_verdict_pack(const unsigned long long* counters, unsigned long long n_err, unsigned long long* v, const FusedSpan* map, uint32_t n_map) {
    unsigned long long unsat = counters[CNT_UNSAT], first = counters[CNT_FIRST_BAD];
    if (n_map) {
        unsat += counters[CNT_FUSED_UNSAT];
        unsigned long long ff = counters[CNT_FUSED_FIRST];
        if (ff != ~0ull) {
            for (uint32_t k = 0; k < n_map; k++)
                if (ff >= map[k].local_base && ff < map[k].local_end) { ff = map[k].global_base + (ff - map[k].local_base); break; }
            first = ff < first ? ff : first;
        }
    }
    v[0] = unsat; v[1] = n_err; v[2] = first; v[3] = counters[CNT_BAD_INPUT];
}

struct Comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    unsigned long long* d_words = nullptr;       // 4 + world words of device scratch
    char err[256] = {0};
    bool active() const { return comm != nullptr; }

    bool fail(const char* what, ncclResult_t r) { snprintf(err, sizeof(err), "%s: %s", what, nccl_api().GetErrorString ? nccl_api().GetErrorString(r) : "NCCL error"); return false; }
    bool init(const uint8_t* id, int rank_, int world_) {
        NcclApi& api = nccl_api();
        if (const char* e = api.load()) { snprintf(err, sizeof(err), "%s", e); return false; }
        ncclUniqueId uid; memcpy(&uid, id, sizeof(uid));
        ncclResult_t r = api.CommInitRank(&comm, world_, uid, rank_);
        if (r != ncclSuccess) { comm = nullptr; return fail("ncclCommInitRank", r); }
        rank = rank_; world = world_;
        if (cudaMalloc(&d_words, (size_t)(8 + 2 * world) * sizeof(unsigned long long)) != cudaSuccess) {
            destroy();                                  // no half-initialised communicator: active() must stay false
            snprintf(err, sizeof(err), "cudaMalloc(comm scratch)"); return false;
        }
        return true;
    }
    void destroy() {
        if (comm) nccl_api().CommDestroy(comm);
        comm = nullptr;
        if (d_words) cudaFree(d_words);
        d_words = nullptr; rank = 0; world = 1;
    }
    // (sum n_unsat, sum n_err, min first_bad, sum bad inputs) over the ranks, from this rank's device counters; returns after the result arrived
    bool allreduce_verdict(const unsigned long long* d_counters, unsigned long long n_err, unsigned long long out[4], const FusedSpan* d_map, uint32_t n_map, cudaStream_t stream) {
        NcclApi& api = nccl_api();
        k_verdict_pack<<<1, 1, 0, stream>>>(d_counters, n_err, d_words, d_map, n_map);
        ncclResult_t r = api.GroupStart(); if (r != ncclSuccess) return fail("ncclGroupStart", r);
        r = api.AllReduce(d_words, d_words, 2, ncclUint64, ncclSum, comm, stream); if (r != ncclSuccess) { api.GroupEnd(); return fail("ncclAllReduce(sum)", r); }
        r = api.AllReduce(d_words + 2, d_words + 2, 1, ncclUint64, ncclMin, comm, stream); if (r != ncclSuccess) { api.GroupEnd(); return fail("ncclAllReduce(min)", r); }
        r = api.AllReduce(d_words + 3, d_words + 3, 1, ncclUint64, ncclSum, comm, stream); if (r != ncclSuccess) { api.GroupEnd(); return fail("ncclAllReduce(sum)", r); }
        r = api.GroupEnd(); if (r != ncclSuccess) return fail("ncclGroupEnd", r);
        if (cudaMemcpyAsync(out, d_words, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) {
            snprintf(err, sizeof(err), "verdict copy: %s", cudaGetErrorString(cudaGetLastError())); return false;
        }
        return true;
    }
    // counts[r] = the value rank r passed (host values; synchronises)
    bool allgather_u64(unsigned long long mine, unsigned long long* counts, cudaStream_t stream) {
        NcclApi& api = nccl_api();
        unsigned long long* d_in = d_words + 4; unsigned long long* d_out = d_words + 8;
        if (cudaMemcpyAsync(d_in, &mine, sizeof(mine), cudaMemcpyHostToDevice, stream) != cudaSuccess) { snprintf(err, sizeof(err), "count upload"); return false; }
        ncclResult_t r = api.AllGather(d_in, d_out, 1, ncclUint64, comm, stream); if (r != ncclSuccess) return fail("ncclAllGather(counts)", r);
        if (cudaMemcpyAsync(counts, d_out, (size_t)world * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream) != cudaSuccess || cudaStreamSynchronize(stream) != cudaSuccess) {
            snprintf(err, sizeof(err), "count copy: %s", cudaGetErrorString(cudaGetLastError())); return false;
        }
        return true;
    }
    // ragged all-gather: rank r's `counts[r]` 32-byte units at `send` land at recv + 32*sum_{j<r} counts[j] on every rank: one
    // broadcast per rank inside a group (NCCL fuses them: every GPU reads the G - 1 other shards over NVLink concurrently)
    bool allgather_ragged(const void* send, void* recv, const unsigned long long* counts, cudaStream_t stream) {
        NcclApi& api = nccl_api();
        // shards of equal length (the even cut of a uniform batch): one ncclAllGather -- NCCL's own all-gather algorithms instead of
        // `world` concurrent broadcasts (PG_GATHER_BCAST=1 forces the general path; tuning runs)
        bool equal = counts[0] != 0;
        for (int g = 1; g < world; g++) equal = equal && counts[g] == counts[0];
        static const bool force_bcast = [] { const char* e = getenv("PG_GATHER_BCAST"); return e && e[0] == '1'; }();
        if (equal && !force_bcast) {
            ncclResult_t ra = api.AllGather(send, recv, (size_t)counts[0] * 4, ncclUint64, comm, stream);
            if (ra != ncclSuccess) return fail("ncclAllGather", ra);
            return true;
        }
        ncclResult_t r = api.GroupStart(); if (r != ncclSuccess) return fail("ncclGroupStart", r);
        unsigned long long off = 0;
        for (int g = 0; g < world; g++) {
            if (counts[g]) {
                r = api.Broadcast(send, (char*)recv + off * 32, (size_t)counts[g] * 4, ncclUint64, g, comm, stream);
                if (r != ncclSuccess) { api.GroupEnd(); return fail("ncclBroadcast", r); }
            }
            off += counts[g];
        }
        r = api.GroupEnd(); if (r != ncclSuccess) return fail("ncclGroupEnd", r);
        return true;
    }
};

}  // namespace pg
