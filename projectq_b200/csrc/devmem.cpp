// Growable device buffer — see devmem.h.
#include "devmem.h"

#include <cuda.h>

#include <new>
#include <stdexcept>
#include <string>

namespace pqb {

namespace {

struct DriverApi {
    bool ok = false;
    CUresult (*memAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*memAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*memCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*memRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*memMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*memUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*memSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*memGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
};

template <class F>
bool fetch(const char* name, F& fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}

const DriverApi& driver() {
    static DriverApi api = [] {
        DriverApi a;
        a.ok = fetch("cuMemAddressReserve", a.memAddressReserve) && fetch("cuMemAddressFree", a.memAddressFree) &&
               fetch("cuMemCreate", a.memCreate) && fetch("cuMemRelease", a.memRelease) && fetch("cuMemMap", a.memMap) &&
               fetch("cuMemUnmap", a.memUnmap) && fetch("cuMemSetAccess", a.memSetAccess) &&
               fetch("cuMemGetAllocationGranularity", a.memGetAllocationGranularity);
        return a;
    }();
    return api;
}

CUmemAllocationProp alloc_prop(int device) {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    return prop;
}

size_t round_up(size_t x, size_t g) { return (x + g - 1) / g * g; }

}  // namespace

void GrowBuffer::init(int device) {
    if (inited_) return;
    device_ = device;
    inited_ = true;
    const DriverApi& d = driver();
    if (!d.ok) return;  // plain cudaMalloc mode
    CUmemAllocationProp prop = alloc_prop(device_);
    size_t gran = 0;
    if (d.memGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0) return;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    const size_t va = round_up(total_b + (size_t(1) << 30), gran);
    CUdeviceptr base = 0;
    if (d.memAddressReserve(&base, va, 0, 0, 0) != CUDA_SUCCESS) return;
    base_ = base;
    va_size_ = va;
    gran_ = gran;
    vmm_ = true;
}

void GrowBuffer::ensure(size_t bytes) {
    if (!inited_) throw std::logic_error("GrowBuffer::ensure before init");
    if (bytes <= mapped_) return;
    if (vmm_) {
        const DriverApi& d = driver();
        if (bytes > va_size_) throw std::bad_alloc();
        const size_t add = round_up(bytes - mapped_, gran_);
        CUmemAllocationProp prop = alloc_prop(device_);
        CUmemGenericAllocationHandle h = 0;
        if (d.memCreate(&h, add, &prop, 0) != CUDA_SUCCESS) throw std::bad_alloc();
        if (d.memMap(base_ + mapped_, add, 0, h, 0) != CUDA_SUCCESS) {
            d.memRelease(h);
            throw std::bad_alloc();
        }
        CUmemAccessDesc acc = {};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = device_;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        if (d.memSetAccess(base_ + mapped_, add, &acc, 1) != CUDA_SUCCESS) {
            d.memUnmap(base_ + mapped_, add);
            d.memRelease(h);
            throw std::bad_alloc();
        }
        chunks_.push_back({h, add});
        mapped_ += add;
        return;
    }
    // cudaMalloc mode: allocate, copy, free
    void* fresh = nullptr;
    if (cudaMalloc(&fresh, bytes) != cudaSuccess) {
        cudaGetLastError();
        throw std::bad_alloc();
    }
    if (mapped_) {
        cudaMemcpy(fresh, reinterpret_cast<void*>(base_), mapped_, cudaMemcpyDeviceToDevice);
        cudaFree(reinterpret_cast<void*>(base_));
    }
    base_ = reinterpret_cast<unsigned long long>(fresh);
    mapped_ = bytes;
}

void GrowBuffer::shrink_to(size_t bytes) {
    if (!inited_) return;
    if (vmm_) {
        const DriverApi& d = driver();
        cudaDeviceSynchronize();
        while (!chunks_.empty() && mapped_ - chunks_.back().size >= bytes) {
            const Chunk ch = chunks_.back();
            chunks_.pop_back();
            mapped_ -= ch.size;
            d.memUnmap(base_ + mapped_, ch.size);
            d.memRelease(ch.handle);
        }
        return;
    }
    if (bytes == 0 && mapped_) {
        cudaDeviceSynchronize();
        cudaFree(reinterpret_cast<void*>(base_));
        base_ = 0;
        mapped_ = 0;
    }
}

GrowBuffer::~GrowBuffer() {
    if (!inited_) return;
    shrink_to(0);
    if (vmm_ && base_) driver().memAddressFree(base_, va_size_);
}

}  // namespace pqb
