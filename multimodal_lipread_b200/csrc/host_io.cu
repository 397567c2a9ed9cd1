// Host-side batch readers for the on-GPU input path: the bytes of a whole batch of `.npy` files
// go straight into a (pinned) ring slot from a pool of native threads, without the interpreter in the loop.
// The reference does this per clip in DataLoader worker processes (np.load + astype(float32) / 255 + permute,
// video/data_utils/dataset_loader.py:87-96); here no arithmetic happens on the host at all.
#include "common.cuh"

#include <atomic>
#include <cerrno>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

namespace hio {

struct NpyHeader {
    std::string descr;
    bool fortran = false;
    std::vector<long long> shape;
    size_t payload = 0;   // file offset of the first data byte
};

static bool read_fully(int fd, void* dst, size_t n, off_t off) {
    char* p = static_cast<char*>(dst);
    while (n) {
        ssize_t r = pread(fd, p, n, off);
        if (r < 0 && errno == EINTR) continue;
        if (r <= 0) return false;
        p += r; off += r; n -= (size_t)r;
    }
    return true;
}

// .npy format 1.0 / 2.0 / 3.0: magic, version, little-endian header length, a Python dict literal.
static const char* parse_header(int fd, NpyHeader& h) {
    unsigned char pre[12];
    if (!read_fully(fd, pre, 10, 0) || memcmp(pre, "\x93NUMPY", 6) != 0) return "not a .npy file";
    size_t hlen, base;
    if (pre[6] == 1) { hlen = pre[8] | (pre[9] << 8); base = 10; }
    else if (pre[6] == 2 || pre[6] == 3) {
        if (!read_fully(fd, pre + 10, 2, 10)) return "truncated header";
        hlen = (size_t)pre[8] | ((size_t)pre[9] << 8) | ((size_t)pre[10] << 16) | ((size_t)pre[11] << 24); base = 12;
    } else return "unsupported .npy version";
    if (hlen > 65536) return "unreasonable header length";
    std::string s(hlen, '\0');
    if (!read_fully(fd, &s[0], hlen, (off_t)base)) return "truncated header";
    h.payload = base + hlen;
    size_t p = s.find("'descr'");
    if (p == std::string::npos) return "no descr in header";
    size_t a = s.find('\'', s.find(':', p)), b = (a == std::string::npos) ? a : s.find('\'', a + 1);
    if (b == std::string::npos) return "structured dtypes are not lip regions / PCM";
    h.descr = s.substr(a + 1, b - a - 1);
    p = s.find("'fortran_order'");
    if (p == std::string::npos) return "no fortran_order in header";
    h.fortran = s.compare(s.find_first_not_of(" :", p + 15), 4, "True") == 0;
    p = s.find("'shape'");
    if (p == std::string::npos) return "no shape in header";
    a = s.find('(', p); b = s.find(')', a);
    if (a == std::string::npos || b == std::string::npos) return "bad shape in header";
    const char* c = s.c_str() + a + 1;
    const char* end = s.c_str() + b;
    while (c < end) {
        while (c < end && (*c == ' ' || *c == ',')) ++c;
        if (c >= end) break;
        char* e;
        long long v = strtoll(c, &e, 10);
        if (e == c) return "bad shape in header";
        h.shape.push_back(v);
        c = e;
    }
    return nullptr;
}

struct Errors {
    std::atomic<int> first{-1};
    std::string msg[1];
    char buf[512] = {0};
    void set(int i, const char* path, const char* what) {
        int expect = -1;
        if (first.compare_exchange_strong(expect, i)) snprintf(buf, sizeof buf, "%s: %s", path, what);
    }
};

template <class F>
static void parallel_for(int n, int n_threads, F&& body) {
    n_threads = n_threads < 1 ? 1 : (n_threads > n ? n : n_threads);
    std::atomic<int> next{0};
    auto work = [&] { for (int i; (i = next.fetch_add(1)) < n;) body(i); };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
}

}  // namespace hio

extern "C" int lr_host_read_npy_u8(const char* const* paths, int n, unsigned char* dst, const long long* shape, int ndim,
                                   int n_threads) {
    LR_CHECK_ARG(n >= 0 && ndim >= 1 && ndim <= 8, "lr_host_read_npy_u8: bad arguments");
    if (n == 0) return LR_OK;
    LR_CHECK_ARG(paths && dst && shape, "lr_host_read_npy_u8: null pointer");
    size_t clip = 1;
    for (int d = 0; d < ndim; ++d) clip *= (size_t)shape[d];
    hio::Errors err;
    hio::parallel_for(n, n_threads, [&](int i) {
        if (err.first.load() >= 0) return;
        int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
        if (fd < 0) { err.set(i, paths[i], strerror(errno)); return; }
        hio::NpyHeader h;
        const char* bad = hio::parse_header(fd, h);
        if (!bad && (h.descr != "|u1" && h.descr != "u1")) bad = "expected uint8 lip regions";
        if (!bad && h.fortran) bad = "expected C-order lip regions";
        if (!bad) {
            bool same = (int)h.shape.size() == ndim;
            for (int d = 0; same && d < ndim; ++d) same = h.shape[d] == shape[d];
            if (!same) bad = "lip regions differ from the batch shape";
        }
        if (!bad && !hio::read_fully(fd, dst + (size_t)i * clip, clip, (off_t)h.payload)) bad = "truncated payload";
        close(fd);
        if (bad) err.set(i, paths[i], bad);
    });
    if (err.first.load() >= 0) return lr::fail(LR_EINVAL, "%s", err.buf);
    return LR_OK;
}

extern "C" int lr_host_read_npy_pcm16(const char* const* paths, int n, short* dst, long long cap, int target_frames,
                                      long long* meta, int n_threads) {
    LR_CHECK_ARG(n >= 0 && cap > 0 && target_frames > 0, "lr_host_read_npy_pcm16: bad arguments");
    if (n == 0) return LR_OK;
    LR_CHECK_ARG(paths && dst && meta, "lr_host_read_npy_pcm16: null pointer");
    hio::Errors err;
    hio::parallel_for(n, n_threads, [&](int i) {
        if (err.first.load() >= 0) return;
        int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
        if (fd < 0) { err.set(i, paths[i], strerror(errno)); return; }
        hio::NpyHeader h;
        const char* bad = hio::parse_header(fd, h);
        if (!bad && h.descr != "<i2") bad = "expected little-endian int16 PCM";
        if (!bad && (h.fortran || h.shape.empty() || h.shape.size() > 2)) bad = "expected C-order PCM of shape (n,) or (n, channels)";
        if (!bad) {
            long long frames = h.shape[0], ch = h.shape.size() == 2 ? h.shape[1] : 1;
            if (frames > target_frames) frames = target_frames;          // later samples are truncated anyway
            if (ch < 1 || ch * (long long)target_frames > cap) bad = "too many channels for the slot";
            else if (frames > 0 && !hio::read_fully(fd, dst + (size_t)i * cap, (size_t)(frames * ch) * 2, (off_t)h.payload))
                bad = "truncated payload";
            else { meta[i] = (long long)i * cap; meta[n + i] = frames; meta[2 * n + i] = ch; }
        }
        close(fd);
        if (bad) err.set(i, paths[i], bad);
    });
    if (err.first.load() >= 0) return lr::fail(LR_EINVAL, "%s", err.buf);
    return LR_OK;
}
