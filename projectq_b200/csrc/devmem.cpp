// Growable device buffer — see devmem.h.
#include "devmem.h"

#include <cuda.h>

#include <unistd.h>

#include <cstdlib>
#include <mutex>
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
    CUresult (*memExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType,
                                           unsigned long long) = nullptr;
    CUresult (*memImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
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
        // optional: sharing a buffer with a partner process
        if (!(fetch("cuMemExportToShareableHandle", a.memExportToShareableHandle) &&
              fetch("cuMemImportFromShareableHandle", a.memImportFromShareableHandle))) {
            a.memExportToShareableHandle = nullptr;
            a.memImportFromShareableHandle = nullptr;
        }
        return a;
    }();
    return api;
}

CUmemAllocationProp alloc_prop(int device, bool exportable) {
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = device;
    // exportable chunks (POSIX file descriptors) only for the shards of a multi-GPU run, where a partner rank maps them
    // for the peer-memory remap; a single-GPU engine allocates exactly as before
    if (exportable) prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    return prop;
}

size_t round_up(size_t x, size_t g) { return (x + g - 1) / g * g; }

// Small mappings are parked instead of torn down: a ProjectQ program creates one Simulator per MainEngine, and reserving
// the address range + mapping the first 64 MB costs ~2 ms of driver calls each time (measured on B200: engine construction
// 1.7-3.7 ms, most of it here).  A destroyed single-GPU buffer that holds exactly its first mapping hands {address range,
// chunk} to this per-process list (at most kMaxParked entries, 64 MB each); init() takes one back when the device matches.
constexpr size_t kMinMapped = size_t(64) << 20;  // 22 qubits
constexpr size_t kMaxParked = 4;
struct Parked {
    int device;
    unsigned long long base, handle;
    size_t va_size, gran, mapped;
};
std::mutex g_parked_mutex;
std::vector<Parked> g_parked;

}  // namespace

void GrowBuffer::init(int device, bool exportable) {
    if (inited_) return;
    device_ = device;
    exportable_ = exportable;
    inited_ = true;
    const DriverApi& d = driver();
    if (!d.ok) return;  // plain cudaMalloc mode
    if (!exportable_) {
        std::lock_guard<std::mutex> lock(g_parked_mutex);
        for (size_t i = 0; i < g_parked.size(); ++i) {
            if (g_parked[i].device != device_) continue;
            const Parked pk = g_parked[i];
            g_parked.erase(g_parked.begin() + long(i));
            base_ = pk.base;
            va_size_ = pk.va_size;
            gran_ = pk.gran;
            mapped_ = pk.mapped;
            chunks_.push_back({pk.handle, pk.mapped});
            vmm_ = true;
            ++generation_;
            return;
        }
    }
    CUmemAllocationProp prop = alloc_prop(device_, exportable_);
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

void GrowBuffer::ensure(size_t bytes, cudaStream_t stream) {
    if (!inited_) throw std::logic_error("GrowBuffer::ensure before init");
    if (bytes <= mapped_) return;
    if (vmm_) {
        const DriverApi& d = driver();
        if (bytes > va_size_) throw std::bad_alloc();
        // Small states grow in one step: every mapping costs three driver calls (cuMemCreate / cuMemMap / cuMemSetAccess,
        // ~1.5 ms together on B200), which dominated allocate_qubit for 20-qubit programs (6.8 ms of an 11 ms QFT-20 replay
        // for the four mappings of 2, 2, 4, 8 MB).  Below kMinMapped the buffer is mapped up to kMinMapped at once.
        size_t want = bytes - mapped_;
        if (bytes < kMinMapped && kMinMapped <= va_size_) want = kMinMapped - mapped_;
        const size_t add = round_up(want, gran_);
        CUmemAllocationProp prop = alloc_prop(device_, exportable_);
        CUmemGenericAllocationHandle h = 0;
        if (d.memCreate(&h, add, &prop, 0) != CUDA_SUCCESS) {
            prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_NONE;  // not exportable here: still usable locally
            if (d.memCreate(&h, add, &prop, 0) != CUDA_SUCCESS) throw std::bad_alloc();
        }
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
        ++generation_;
        return;
    }
    // cudaMalloc mode: allocate, copy, free
    void* fresh = nullptr;
    if (cudaMalloc(&fresh, bytes) != cudaSuccess) {
        cudaGetLastError();
        throw std::bad_alloc();
    }
    if (mapped_) {
        // the engine's stream is non-blocking, so a legacy-stream copy would not wait for kernels still in flight on it
        cudaError_t e = cudaMemcpyAsync(fresh, reinterpret_cast<void*>(base_), mapped_, cudaMemcpyDeviceToDevice, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaFree(fresh);
            throw std::runtime_error(std::string("GrowBuffer: copy while growing failed: ") + cudaGetErrorString(e));
        }
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
            ++generation_;
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

bool GrowBuffer::export_chunks(std::vector<int>& fds, std::vector<size_t>& sizes) const {
    fds.clear();
    sizes.clear();
    const DriverApi& d = driver();
    if (!vmm_ || !d.memExportToShareableHandle) return false;
    for (auto& ch : chunks_) {
        int fd = -1;
        if (d.memExportToShareableHandle(&fd, ch.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) != CUDA_SUCCESS || fd < 0) {
            for (int f : fds) ::close(f);
            fds.clear();
            sizes.clear();
            return false;
        }
        fds.push_back(fd);
        sizes.push_back(ch.size);
    }
    return true;
}

void PeerMapping::map(int device, const std::vector<int>& fds, const std::vector<size_t>& sizes) {
    reset();
    const DriverApi& d = driver();
    if (!d.ok || !d.memImportFromShareableHandle) throw std::runtime_error("peer mapping: driver API unavailable");
    size_t total = 0;
    for (auto s : sizes) total += s;
    CUdeviceptr base = 0;
    if (d.memAddressReserve(&base, total, 0, 0, 0) != CUDA_SUCCESS) throw std::runtime_error("peer mapping: reserve failed");
    base_ = base;
    va_size_ = total;
    size_t off = 0;
    for (size_t i = 0; i < fds.size(); ++i) {
        CUmemGenericAllocationHandle h = 0;
        if (d.memImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fds[i])),
                                           CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS) {
            reset();
            throw std::runtime_error("peer mapping: import failed");
        }
        if (d.memMap(base_ + off, sizes[i], 0, h, 0) != CUDA_SUCCESS) {
            d.memRelease(h);
            reset();
            throw std::runtime_error("peer mapping: map failed");
        }
        handles_.emplace_back(h, sizes[i]);
        off += sizes[i];
        total_ = off;
    }
    CUmemAccessDesc acc = {};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    if (d.memSetAccess(base_, total_, &acc, 1) != CUDA_SUCCESS) {
        reset();
        throw std::runtime_error("peer mapping: no peer access between the two devices");
    }
}

void PeerMapping::reset() {
    if (!base_) return;
    const DriverApi& d = driver();
    cudaDeviceSynchronize();
    size_t off = 0;
    for (auto& h : handles_) {
        d.memUnmap(base_ + off, h.second);
        d.memRelease(h.first);
        off += h.second;
    }
    handles_.clear();
    d.memAddressFree(base_, va_size_);
    base_ = 0;
    total_ = va_size_ = 0;
}

GrowBuffer::~GrowBuffer() {
    if (!inited_) return;
    if (vmm_ && !exportable_ && base_ && !chunks_.empty()) {
        shrink_to(1);  // back to the first mapping (waits for the device)
        if (chunks_.size() == 1 && chunks_[0].size == mapped_ && mapped_ <= kMinMapped) {
            std::lock_guard<std::mutex> lock(g_parked_mutex);
            if (g_parked.size() < kMaxParked) {
                g_parked.push_back({device_, base_, chunks_[0].handle, va_size_, gran_, mapped_});
                chunks_.clear();
                base_ = 0;
                mapped_ = 0;
                return;
            }
        }
    }
    shrink_to(0);
    if (vmm_ && base_) driver().memAddressFree(base_, va_size_);
}

}  // namespace pqb
