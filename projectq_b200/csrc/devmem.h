// Growable device buffer for the state vector.
//
// The reference grows the state by allocating a vector of twice the size and copying (reference:
// simulator.hpp:55-74, with the static tmpBuff1_/tmpBuff2_ recycling at :574-578).  On the GPU the state may use most of
// the 180 GB of HBM, so a copy-on-grow peak of 1.5x is not affordable: the buffer reserves a virtual address range once
// and maps physical memory behind it as the state doubles (CUDA virtual memory management), so growing never copies.
// The driver entry points are fetched through the runtime (cudaGetDriverEntryPoint) so the library does not link
// libcuda and still loads on a machine without a driver (CPU-side tests of the host logic).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <vector>

namespace pqb {

class GrowBuffer {
public:
    GrowBuffer() = default;
    ~GrowBuffer();
    GrowBuffer(const GrowBuffer&) = delete;
    GrowBuffer& operator=(const GrowBuffer&) = delete;

    void init(int device, bool exportable = false);
    // make at least `bytes` usable; existing contents are preserved; throws std::bad_alloc when the device is full.
    // `stream` = the stream whose queued work may still be writing the buffer: the (non-VMM fallback) grow-by-copy is
    // ordered after it.
    void ensure(size_t bytes, cudaStream_t stream = nullptr);
    // give physical memory back, keeping at least `bytes` mapped
    void shrink_to(size_t bytes);
    void release() { shrink_to(0); }

    // Export every mapped chunk as a POSIX file descriptor (caller closes them) so that a partner process can map the
    // buffer (PeerMapping).  Returns false when the buffer is not a VMM allocation or the driver refuses.
    bool export_chunks(std::vector<int>& fds, std::vector<size_t>& sizes) const;
    // changes whenever the set of chunks changes (a partner's mapping is then stale); never repeats within a process
    uint64_t layout_key() const { return (base_ >> 12) * 0x9E3779B97F4A7C15ULL + generation_ + 1; }

    void* ptr() const { return reinterpret_cast<void*>(base_); }
    double2* amps() const { return reinterpret_cast<double2*>(base_); }
    size_t capacity() const { return mapped_; }
    bool uses_vmm() const { return vmm_; }

private:
    struct Chunk {
        unsigned long long handle;
        size_t size;
    };
    int device_ = 0;
    bool vmm_ = false;
    bool exportable_ = false;
    bool inited_ = false;
    unsigned long long base_ = 0;  // CUdeviceptr
    size_t va_size_ = 0;
    size_t mapped_ = 0;
    size_t gran_ = 0;
    uint64_t generation_ = 0;  // bumped by every change of the chunk list
    std::vector<Chunk> chunks_;
};

// A partner rank's buffer mapped into this process (peer access over NVLink) from the descriptors of its chunks.
class PeerMapping {
public:
    PeerMapping() = default;
    ~PeerMapping() { reset(); }
    PeerMapping(const PeerMapping&) = delete;
    PeerMapping& operator=(const PeerMapping&) = delete;
    // import the chunks (descriptors stay owned by the caller), map them back to back, grant `device` read/write access
    void map(int device, const std::vector<int>& fds, const std::vector<size_t>& sizes);
    void reset();
    double2* amps() const { return reinterpret_cast<double2*>(base_); }
    size_t bytes() const { return total_; }

private:
    unsigned long long base_ = 0;
    size_t total_ = 0, va_size_ = 0;
    std::vector<std::pair<unsigned long long, size_t>> handles_;
};

}  // namespace pqb
