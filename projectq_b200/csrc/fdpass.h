// Passing file descriptors between the rank processes of one box (Unix-domain sockets, SCM_RIGHTS).
//
// The shards are CUDA virtual-memory allocations (devmem.h); a partner rank can map one into its own address space only
// from an exported POSIX file descriptor, and a descriptor has to travel through a Unix socket.  Every rank listens on an
// abstract socket named after the run (a tag shared by all ranks, e.g. a hash of the NCCL id) and its rank; the
// higher rank of a pair connects, the lower one accepts.  Pure host code, unit-tested on the CPU (pqb_host_fdpass_selftest).
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace pqb {

class FdChannel {
public:
    FdChannel(uint64_t run_tag, int rank, int world);
    ~FdChannel();
    FdChannel(const FdChannel&) = delete;
    FdChannel& operator=(const FdChannel&) = delete;

    // send / receive one message: a small payload plus any number of descriptors (received descriptors are owned by
    // the caller, who must close them)
    void send(int peer, const void* payload, size_t n_bytes, const std::vector<int>& fds);
    void recv(int peer, void* payload, size_t n_bytes, std::vector<int>& fds, size_t n_fds);

private:
    int socket_to(int peer);  // connects (higher rank) or accepts (lower rank) on first use
    std::string name_of(int rank) const;
    uint64_t tag_;
    int rank_, world_;
    int listen_fd_ = -1;
    std::map<int, int> conn_;  // peer -> connected socket
};

}  // namespace pqb
