// ssb_fabric.cuh -- one address space over the GPUs of a box (x-slab sharding).
//
// A "stitched" array is a virtual range of world * chunk bytes whose p-th chunk is physical memory
// of rank p's GPU (cuMemCreate, exported as a POSIX file descriptor; the host side passes the
// descriptors between the ranks -- one process per GPU -- over unix sockets), mapped on every rank.  Kernels index it like an ordinary array: the part of
// the map / the agent slots a rank owns is local HBM, the rest is reached over NVLink (loads,
// stores and atomics), so the halo of the slab decomposition needs no packing or exchange step.
// The driver entry points are resolved at run time (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>
#include <vector>
#include "../../include/ssb.h"

struct ssb_fabric {
    int rank, world, device;
    size_t granularity;
    struct Array {
        size_t chunk;                              // this rank's chunk
        size_t size[SSB_MAX_RANKS], total;         // all chunks, once mapped
        CUmemGenericAllocationHandle part[SSB_MAX_RANKS];
        bool have[SSB_MAX_RANKS];
        CUdeviceptr base;
        bool mapped;
    };
    std::vector<Array> arrays;
    char err[256];
    // driver entry points
    CUresult (*memCreate)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long);
    CUresult (*memRelease)(CUmemGenericAllocationHandle);
    CUresult (*memGetGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags);
    CUresult (*memExport)(void *, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
    CUresult (*memImport)(CUmemGenericAllocationHandle *, void *, CUmemAllocationHandleType);
    CUresult (*memAddressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long);
    CUresult (*memAddressFree)(CUdeviceptr, size_t);
    CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
    CUresult (*memUnmap)(CUdeviceptr, size_t);
    CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t);
};

namespace ssb {

template <class F>
static bool fabric_entry(const char *name, F *out) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return false;
    *out = (F)fn;
    return true;
}

static int fabric_fail(ssb_fabric *f, const char *what, CUresult rc) {
    snprintf(f->err, sizeof f->err, "%s failed (CUresult %d)", what, (int)rc);
    return -1;
}

static CUmemAllocationProp fabric_prop(int device) {
    CUmemAllocationProp prop;
    memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

}  // namespace ssb

extern "C" {

const char *ssb_fabric_last_error(ssb_fabric_t *f) { return f ? f->err : "null fabric"; }

int ssb_fabric_create(int32_t rank, int32_t world, int32_t device, ssb_fabric_t **out) {
    using namespace ssb;
    if (!out || world < 1 || world > SSB_MAX_RANKS || rank < 0 || rank >= world) return -1;
    ssb_fabric *f = new ssb_fabric();
    f->err[0] = 0;
    f->rank = rank; f->world = world; f->device = device;
    *out = f;
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(0) != cudaSuccess) { snprintf(f->err, sizeof f->err, "cudaSetDevice(%d) failed", device); return -1; }
    const bool ok = fabric_entry("cuMemCreate", &f->memCreate) && fabric_entry("cuMemRelease", &f->memRelease) &&
                    fabric_entry("cuMemGetAllocationGranularity", &f->memGetGranularity) &&
                    fabric_entry("cuMemExportToShareableHandle", &f->memExport) &&
                    fabric_entry("cuMemImportFromShareableHandle", &f->memImport) &&
                    fabric_entry("cuMemAddressReserve", &f->memAddressReserve) && fabric_entry("cuMemAddressFree", &f->memAddressFree) &&
                    fabric_entry("cuMemMap", &f->memMap) && fabric_entry("cuMemUnmap", &f->memUnmap) &&
                    fabric_entry("cuMemSetAccess", &f->memSetAccess);
    if (!ok) { snprintf(f->err, sizeof f->err, "CUDA virtual memory management entry points not available"); return -1; }
    CUmemAllocationProp prop = fabric_prop(device);
    CUresult rc = f->memGetGranularity(&f->granularity, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    if (rc != CUDA_SUCCESS) return fabric_fail(f, "cuMemGetAllocationGranularity", rc);
    return 0;
}

uint64_t ssb_fabric_granularity(ssb_fabric_t *f) { return f ? f->granularity : 0; }

// This rank's chunk of the next stitched array: physical memory on this GPU, exported as a file
// descriptor for the peers.  Returns the array's index.  (cuMemMap cannot map a sub-range of an
// allocation, so every array has its own allocation per rank.)
int ssb_fabric_alloc(ssb_fabric_t *f, uint64_t chunk_bytes, int32_t *fd_out) {
    using namespace ssb;
    if (!f || !fd_out || chunk_bytes == 0) return -1;
    ssb_fabric::Array a;
    memset(&a, 0, sizeof a);
    a.chunk = (chunk_bytes + f->granularity - 1) / f->granularity * f->granularity;
    CUmemAllocationProp prop = fabric_prop(f->device);
    CUresult rc = f->memCreate(&a.part[f->rank], a.chunk, &prop, 0);
    if (rc != CUDA_SUCCESS) return fabric_fail(f, "cuMemCreate", rc);
    a.have[f->rank] = true;
    int fd = -1;
    rc = f->memExport(&fd, a.part[f->rank], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
    if (rc != CUDA_SUCCESS) return fabric_fail(f, "cuMemExportToShareableHandle", rc);
    *fd_out = fd;
    f->arrays.push_back(a);
    return (int)f->arrays.size() - 1;
}

// Import the peers' chunks of array `index` (fds[p] for every p != rank; closed here) and map the
// stitched array: base[p * chunk ..] lives on rank p.
int ssb_fabric_map(ssb_fabric_t *f, int32_t index, const int32_t *fds, const uint64_t *chunk_bytes, void **base_out) {
    using namespace ssb;
    if (!f || !base_out || !chunk_bytes || index < 0 || index >= (int)f->arrays.size() || (f->world > 1 && !fds)) return -1;
    ssb_fabric::Array &a = f->arrays[index];
    if (a.mapped) { snprintf(f->err, sizeof f->err, "array %d already mapped", index); return -1; }
    a.total = 0;
    for (int p = 0; p < f->world; p++) {
        if (chunk_bytes[p] == 0 || chunk_bytes[p] % f->granularity) { snprintf(f->err, sizeof f->err, "chunk %d of array %d: %llu bytes is not a multiple of the granularity", p, index, (unsigned long long)chunk_bytes[p]); return -1; }
        a.size[p] = chunk_bytes[p];
        a.total += chunk_bytes[p];
    }
    if (a.size[f->rank] != a.chunk) { snprintf(f->err, sizeof f->err, "array %d: own chunk size mismatch", index); return -1; }
    CUresult rc;
    // error paths give everything back: the peers' descriptors not yet imported are closed, imported handles released, the
    // parts already mapped unmapped and the address range freed
    int next_fd = 0;
    size_t mapped = 0;
    bool reserved = false;
    auto undo = [&](const char *what, CUresult code) {
        for (int p = next_fd; p < f->world; p++) if (p != f->rank) close(fds[p]);
        if (mapped) f->memUnmap(a.base, mapped);
        if (reserved) f->memAddressFree(a.base, a.total);
        for (int p = 0; p < f->world; p++) if (p != f->rank && a.have[p]) { f->memRelease(a.part[p]); a.have[p] = false; }
        return fabric_fail(f, what, code);
    };
    for (int p = 0; p < f->world; p++) {
        next_fd = p;
        if (p == f->rank) continue;
        rc = f->memImport(&a.part[p], (void *)(uintptr_t)fds[p], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        if (rc != CUDA_SUCCESS) return undo("cuMemImportFromShareableHandle", rc);
        a.have[p] = true;
        close(fds[p]);
    }
    next_fd = f->world;
    rc = f->memAddressReserve(&a.base, a.total, f->granularity, 0, 0);
    if (rc != CUDA_SUCCESS) return undo("cuMemAddressReserve", rc);
    reserved = true;
    for (int p = 0; p < f->world; p++) {
        rc = f->memMap(a.base + mapped, a.size[p], 0, a.part[p], 0);
        if (rc != CUDA_SUCCESS) return undo("cuMemMap", rc);
        mapped += a.size[p];
    }
    CUmemAccessDesc acc;
    memset(&acc, 0, sizeof acc);
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = f->device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    rc = f->memSetAccess(a.base, a.total, &acc, 1);
    if (rc != CUDA_SUCCESS) return undo("cuMemSetAccess", rc);
    a.mapped = true;
    *base_out = (void *)a.base;
    return 0;
}

int ssb_fabric_copy(ssb_fabric_t *f, void *dst, const void *src, uint64_t bytes, int32_t kind) {
    if (!f || !dst || (kind != 2 && !src)) return -1;
    cudaError_t rc = cudaSetDevice(f->device);
    if (rc == cudaSuccess) {
        if (kind == 0) rc = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
        else if (kind == 1) rc = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
        else rc = cudaMemset(dst, 0, bytes);
    }
    if (rc == cudaSuccess) rc = cudaDeviceSynchronize();
    if (rc != cudaSuccess) { snprintf(f->err, sizeof f->err, "fabric copy (kind %d, %llu bytes) failed: %s", kind, (unsigned long long)bytes, cudaGetErrorString(rc)); return -2; }
    return 0;
}

int ssb_fabric_destroy(ssb_fabric_t *f) {
    if (!f) return 0;
    cudaSetDevice(f->device);
    cudaDeviceSynchronize();
    for (auto &a : f->arrays) {
        if (a.mapped) {
            f->memUnmap(a.base, a.total);
            f->memAddressFree(a.base, a.total);
        }
        for (int p = 0; p < f->world; p++) if (a.have[p]) f->memRelease(a.part[p]);
    }
    delete f;
    return 0;
}

}  // extern "C"
