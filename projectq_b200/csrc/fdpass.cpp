// Descriptor passing between rank processes — see fdpass.h.
#include "fdpass.h"

#include <sys/socket.h>
#include <sys/time.h>
#include <sys/un.h>
#include <unistd.h>

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <thread>

namespace pqb {

namespace {

void fail(const char* what) { throw std::runtime_error(std::string("fdpass: ") + what + ": " + std::strerror(errno)); }

// abstract-namespace address (leading NUL): no file system entry, vanishes with the process
socklen_t make_addr(const std::string& name, sockaddr_un& addr) {
    std::memset(&addr, 0, sizeof(addr));
    addr.sun_family = AF_UNIX;
    if (name.size() + 1 > sizeof(addr.sun_path)) throw std::runtime_error("fdpass: socket name too long");
    std::memcpy(addr.sun_path + 1, name.data(), name.size());
    return socklen_t(offsetof(sockaddr_un, sun_path) + 1 + name.size());
}

void write_all(int fd, const void* p, size_t n) {
    const char* c = static_cast<const char*>(p);
    while (n) {
        const ssize_t w = ::send(fd, c, n, MSG_NOSIGNAL);
        if (w < 0) {
            if (errno == EINTR) continue;
            fail("send");
        }
        c += w;
        n -= size_t(w);
    }
}

void read_all(int fd, void* p, size_t n) {
    char* c = static_cast<char*>(p);
    while (n) {
        const ssize_t r = ::recv(fd, c, n, 0);
        if (r < 0) {
            if (errno == EINTR) continue;
            if (errno == EAGAIN || errno == EWOULDBLOCK) throw std::runtime_error("fdpass: the partner rank did not answer in time");
            fail("recv");
        }
        if (r == 0) throw std::runtime_error("fdpass: peer closed the connection");
        c += r;
        n -= size_t(r);
    }
}

constexpr size_t kMaxFdsPerMsg = 32;

// a partner that died must not leave this rank blocked forever: give every connected socket a receive time limit
void set_receive_timeout(int fd) {
    timeval tv;
    tv.tv_sec = 300;
    tv.tv_usec = 0;
    ::setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof(tv));
}

}  // namespace

std::string FdChannel::name_of(int rank) const {
    char buf[64];
    std::snprintf(buf, sizeof(buf), "pqb200-%016llx-%d", static_cast<unsigned long long>(tag_), rank);
    return buf;
}

FdChannel::FdChannel(uint64_t run_tag, int rank, int world) : tag_(run_tag), rank_(rank), world_(world) {
    listen_fd_ = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (listen_fd_ < 0) fail("socket");
    sockaddr_un addr;
    const socklen_t len = make_addr(name_of(rank_), addr);
    if (::bind(listen_fd_, reinterpret_cast<sockaddr*>(&addr), len) < 0) fail("bind");
    if (::listen(listen_fd_, 64) < 0) fail("listen");
}

FdChannel::~FdChannel() {
    for (auto& kv : conn_) ::close(kv.second);
    if (listen_fd_ >= 0) ::close(listen_fd_);
}

int FdChannel::socket_to(int peer) {
    auto it = conn_.find(peer);
    if (it != conn_.end()) return it->second;
    if (peer == rank_) throw std::runtime_error("fdpass: no channel to self");
    if (rank_ > peer) {
        // the higher rank connects; the listener may not be up yet if the partner is still starting
        sockaddr_un addr;
        const socklen_t len = make_addr(name_of(peer), addr);
        int fd = -1;
        for (int attempt = 0; attempt < 3000; ++attempt) {
            fd = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
            if (fd < 0) fail("socket");
            if (::connect(fd, reinterpret_cast<sockaddr*>(&addr), len) == 0) break;
            ::close(fd);
            fd = -1;
            if (errno != ECONNREFUSED && errno != ENOENT && errno != EAGAIN) fail("connect");
            std::this_thread::sleep_for(std::chrono::milliseconds(10));
        }
        if (fd < 0) throw std::runtime_error("fdpass: partner rank is not listening");
        set_receive_timeout(fd);
        const int32_t me = rank_;
        write_all(fd, &me, sizeof(me));
        conn_[peer] = fd;
        return fd;
    }
    // the lower rank accepts until the expected partner shows up (others that arrive early are kept)
    while (true) {
        const int fd = ::accept4(listen_fd_, nullptr, nullptr, SOCK_CLOEXEC);
        if (fd < 0) {
            if (errno == EINTR) continue;
            fail("accept");
        }
        // only processes of the same user may join: the descriptors passed here give read/write access to the GPU state
        ucred cred;
        socklen_t clen = sizeof(cred);
        if (::getsockopt(fd, SOL_SOCKET, SO_PEERCRED, &cred, &clen) != 0 || cred.uid != ::geteuid()) {
            ::close(fd);
            continue;
        }
        set_receive_timeout(fd);
        int32_t who = -1;
        read_all(fd, &who, sizeof(who));
        // a rank connects only to lower ranks, and only once
        if (who <= rank_ || who >= world_ || conn_.count(who)) {
            ::close(fd);
            continue;
        }
        conn_[who] = fd;
        if (who == peer) return fd;
    }
}

void FdChannel::send(int peer, const void* payload, size_t n_bytes, const std::vector<int>& fds) {
    const int s = socket_to(peer);
    if (fds.size() > kMaxFdsPerMsg) throw std::runtime_error("fdpass: too many descriptors in one message");
    // header: payload size and descriptor count, with the descriptors attached to it
    uint64_t header[2] = {uint64_t(n_bytes), uint64_t(fds.size())};
    msghdr msg;
    std::memset(&msg, 0, sizeof(msg));
    iovec iov;
    iov.iov_base = header;
    iov.iov_len = sizeof(header);
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    alignas(cmsghdr) char control[CMSG_SPACE(sizeof(int) * kMaxFdsPerMsg)];
    if (!fds.empty()) {
        std::memset(control, 0, sizeof(control));
        msg.msg_control = control;
        msg.msg_controllen = CMSG_SPACE(sizeof(int) * fds.size());
        cmsghdr* c = CMSG_FIRSTHDR(&msg);
        c->cmsg_level = SOL_SOCKET;
        c->cmsg_type = SCM_RIGHTS;
        c->cmsg_len = CMSG_LEN(sizeof(int) * fds.size());
        std::memcpy(CMSG_DATA(c), fds.data(), sizeof(int) * fds.size());
    }
    while (true) {
        const ssize_t w = ::sendmsg(s, &msg, MSG_NOSIGNAL);
        if (w == ssize_t(sizeof(header))) break;
        if (w < 0 && errno == EINTR) continue;
        fail("sendmsg");
    }
    if (n_bytes) write_all(s, payload, n_bytes);
}

void FdChannel::recv(int peer, void* payload, size_t n_bytes, std::vector<int>& fds, size_t n_fds) {
    const int s = socket_to(peer);
    uint64_t header[2] = {0, 0};
    msghdr msg;
    std::memset(&msg, 0, sizeof(msg));
    iovec iov;
    iov.iov_base = header;
    iov.iov_len = sizeof(header);
    msg.msg_iov = &iov;
    msg.msg_iovlen = 1;
    alignas(cmsghdr) char control[CMSG_SPACE(sizeof(int) * kMaxFdsPerMsg)];
    std::memset(control, 0, sizeof(control));
    msg.msg_control = control;
    msg.msg_controllen = sizeof(control);
    while (true) {
        const ssize_t r = ::recvmsg(s, &msg, MSG_CMSG_CLOEXEC | MSG_WAITALL);
        if (r == ssize_t(sizeof(header))) break;
        if (r < 0 && errno == EINTR) continue;
        if (r == 0) throw std::runtime_error("fdpass: peer closed the connection");
        if (r < 0 && (errno == EAGAIN || errno == EWOULDBLOCK))
            throw std::runtime_error("fdpass: the partner rank did not answer in time");
        fail("recvmsg");
    }
    fds.clear();
    for (cmsghdr* c = CMSG_FIRSTHDR(&msg); c; c = CMSG_NXTHDR(&msg, c)) {
        if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
            const size_t n = (c->cmsg_len - CMSG_LEN(0)) / sizeof(int);
            const int* p = reinterpret_cast<const int*>(CMSG_DATA(c));
            fds.insert(fds.end(), p, p + n);
        }
    }
    if (header[0] != n_bytes || header[1] != n_fds || fds.size() != n_fds) {
        for (int fd : fds) ::close(fd);
        fds.clear();
        throw std::runtime_error("fdpass: unexpected message shape from the partner rank");
    }
    if (n_bytes) read_all(s, payload, n_bytes);
}

}  // namespace pqb
