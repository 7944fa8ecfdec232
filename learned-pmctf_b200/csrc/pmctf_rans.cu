// pmctf_rans.cu -- the entropy-coder boundary of pWave++ / pMCTF.
//
// What it replaces in the reference:
//   pMCTF/cpp/rans/rans.cpp:76-168,272-331      RansEncoderLib / RansDecoderLib (64-bit rANS, 16-bit probabilities, 4-bit
//                                               bypass digits for out-of-range symbols)
//   pMCTF/cpp/py_rans/py_rans.cpp:22-225        RansEncoder / RansDecoder: symbol array split into `streamPart` sub-streams,
//                                               one flag byte + per-stream sizes in front of the concatenated streams
//   pMCTF/cpp/ops/ops.cpp:24-82                 pmf_to_quantized_cdf
//   pMCTF/entropy_models/entropy_models.py:37-40,266-275   the per-call `.clamp().to(int16).cpu()` of the symbols and of the
//                                               scale indexes (build_indexes), one blocking device->host copy each
//
// The rANS core of the reference is rygorous/ryg_rans `rans64.h` (pinned at c9d162d996fd600315af9ae8eb89d832576cb32d by
// cpp/3rdparty/ryg_rans/CMakeLists.txt.in, NOT vendored in the reference tree).  Its published algorithm is restated here
// (state in [2^31, 2^63), 32-bit renormalisation words written backwards, x' = (x / f << bits) + x % f + start):
// streams are byte-identical to what the reference's encoder produces (tests/test_rans.py checks that against the
// reference's own rans.cpp compiled against the restated header, oracle/_ref).
//
// Split of the work on a B200 node: the GPU turns the fp32 symbol / scale planes into int16 symbols + int16 table indexes in
// ONE pass (gaussian_symbolize_kernel, HBM-bound, 12 B per coefficient) so a single copy per coded step reaches the host;
// the strictly sequential state update runs on the host cores, one thread per sub-stream, behind a C ABI.
#include <cuda_runtime.h>

#include <algorithm>
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "pmctf_b200.h"

namespace pmctf {
void count_launch();   // pmctf_kernels.cu

namespace rans {

constexpr uint64_t LOWER = 1ull << 31;          // lower bound of the normalised state interval
constexpr uint32_t PROB_BITS = 16;              // probability resolution of the CDF tables
constexpr uint32_t RAW_BITS = 4;                // width of one bypass digit
constexpr uint32_t RAW_MAX = (1u << RAW_BITS) - 1;

// One queued coding decision: a table interval [start, start + range) of 2^16, or (range == 0) a raw 4-bit digit `start`.
struct Interval {
    uint16_t start, range;
};

struct Writer {   // 32-bit words written from the end of a buffer towards its start
    uint32_t *p;
    uint64_t x = LOWER;
    inline void put(uint32_t start, uint32_t freq, uint32_t bits)
    {
        const uint64_t limit = ((LOWER >> bits) << 32) * freq;
        if (x >= limit) {
            *--p = (uint32_t)x;
            x >>= 32;
        }
        x = ((x / freq) << bits) + (x % freq) + start;
    }
    inline void put_raw(uint32_t val, uint32_t nbits)
    {   // a uniform symbol of 2^nbits values: freq = 2^(16 - nbits) at 16-bit resolution
        const uint64_t limit = ((LOWER >> 16) << 32) * (uint64_t)(1u << (16 - nbits));
        if (x >= limit) {
            *--p = (uint32_t)x;
            x >>= 32;
        }
        x = (x << nbits) | val;
    }
    inline void finish()
    {
        p -= 2;
        p[0] = (uint32_t)x;
        p[1] = (uint32_t)(x >> 32);
    }
};

struct Reader {
    const uint32_t *p = nullptr, *end = nullptr;
    uint64_t x = 0;
    bool overrun = false;
    inline uint32_t next()
    {
        if (p >= end) {
            overrun = true;
            return 0;
        }
        return *p++;
    }
    void start(const uint32_t *words, size_t n)
    {
        p = words;
        end = words + n;
        overrun = false;
        const uint64_t lo = next(), hi = next();
        x = lo | (hi << 32);
    }
    inline uint32_t peek(uint32_t bits) const { return (uint32_t)(x & ((1ull << bits) - 1)); }
    inline void advance(uint32_t start, uint32_t freq, uint32_t bits)
    {
        x = (uint64_t)freq * (x >> bits) + (x & ((1ull << bits) - 1)) - start;
        if (x < LOWER) x = (x << 32) | next();
    }
    inline uint32_t get_raw(uint32_t nbits)
    {
        const uint32_t v = (uint32_t)(x & ((1u << nbits) - 1));
        x >>= nbits;
        if (x < LOWER) x = (x << 32) | next();
        return v;
    }
};

struct Tables {   // one shared copy of the caller's CDF tables per encode / decode call
    std::vector<int32_t> cdf;      // [num][stride]
    std::vector<int32_t> sizes, offsets;
    int num = 0, stride = 0;
    bool monotone = true;          // every row non-decreasing over its size: the decoder may bisect
};

// queue the intervals of `n` symbols (rans.cpp:76-139): symbols whose table index is negative are skipped, values outside
// the table go through the table's last entry (the escape) followed by raw digits: count of digits in unary-ish base 15,
// then the digits of the zig-zag folded remainder, least significant first
static int queue_symbols(std::vector<Interval> &q, const int16_t *sym, const int16_t *idx, long long n, const Tables &t)
{
    q.reserve(q.size() + (size_t)n + (size_t)n / 2);
    for (long long i = 0; i < n; ++i) {
        const int ti = idx[i];
        if (ti < 0) continue;
        if (ti >= t.num) return PMCTF_EINVAL;
        const int32_t *cdf = t.cdf.data() + (size_t)ti * t.stride;
        const int32_t escape = t.sizes[ti] - 2;
        if (escape < 0 || escape + 1 >= t.stride) return PMCTF_EINVAL;
        int32_t v = (int32_t)sym[i] - t.offsets[ti];
        uint32_t raw = 0;
        if (v < 0) {
            raw = (uint32_t)(-2 * v - 1);
            v = escape;
        } else if (v >= escape) {
            raw = (uint32_t)(2 * (v - escape));
            v = escape;
        }
        const uint16_t range = (uint16_t)(cdf[v + 1] - cdf[v]);
        if (range == 0) return PMCTF_EINVAL;   // an interval of zero width cannot be coded
        q.push_back({(uint16_t)cdf[v], range});
        if (v == escape) {
            int digits = 0;
            while ((raw >> (digits * RAW_BITS)) != 0) ++digits;
            int c = digits;
            while (c >= (int)RAW_MAX) {
                q.push_back({(uint16_t)RAW_MAX, 0});
                c -= RAW_MAX;
            }
            q.push_back({(uint16_t)c, 0});
            for (int j = 0; j < digits; ++j) q.push_back({(uint16_t)((raw >> (j * RAW_BITS)) & RAW_MAX), 0});
        }
    }
    return 0;
}

// rans.cpp:141-168: the queue is coded last-in first-out so the decoder reads the symbols in their original order
static void code_queue(std::vector<Interval> &q, std::vector<uint8_t> &bytes, std::vector<uint32_t> &words)
{
    if (words.size() < q.size() + 2) words.resize(q.size() + 2);   // grow-only scratch of the sub-encoder: no fresh pages per flush
    Writer w;
    w.p = words.data() + words.size();
    for (size_t i = q.size(); i-- > 0;) {
        const Interval s = q[i];
        if (s.range)
            w.put(s.start, s.range, PROB_BITS);
        else
            w.put_raw(s.start, RAW_BITS);
    }
    w.finish();
    q.clear();
    const size_t n = (size_t)(words.data() + words.size() - w.p) * sizeof(uint32_t);
    bytes.resize(n);
    memcpy(bytes.data(), w.p, n);
}

// A sub-stream of the encoder.  With a worker thread the calls return at once and run in order on that thread
// (RansEncoderLibMultiThread, rans.cpp:170-270); stream() waits for the flush that was queued before it.
class SubEncoder {
public:
    explicit SubEncoder(bool threaded) : threaded_(threaded)
    {
        if (threaded_) worker_ = std::thread([this] { run(); });
    }
    ~SubEncoder()
    {
        if (threaded_) {
            {
                std::lock_guard<std::mutex> g(m_);
                stop_ = true;
            }
            cv_.notify_all();
            worker_.join();
        }
    }
    void submit(std::function<void()> f)
    {
        if (!threaded_) {
            f();
            return;
        }
        {
            std::lock_guard<std::mutex> g(m_);
            jobs_.push_back(std::move(f));
            ++pending_;
        }
        cv_.notify_all();
    }
    void drain()
    {
        if (!threaded_) return;
        std::unique_lock<std::mutex> g(m_);
        done_.wait(g, [this] { return pending_ == 0; });
    }
    std::vector<Interval> queue;
    std::vector<uint8_t> bytes;
    std::vector<uint32_t> words;
    int error = 0;

private:
    void run()
    {
        std::unique_lock<std::mutex> g(m_);
        for (;;) {
            cv_.wait(g, [this] { return stop_ || !jobs_.empty(); });
            if (jobs_.empty()) return;   // stop requested and nothing left
            std::function<void()> f = std::move(jobs_.front());
            jobs_.pop_front();
            g.unlock();
            f();
            g.lock();
            if (--pending_ == 0) done_.notify_all();
        }
    }
    bool threaded_;
    bool stop_ = false;
    int pending_ = 0;
    std::thread worker_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::deque<std::function<void()>> jobs_;
};

struct Encoder {
    std::vector<std::unique_ptr<SubEncoder>> parts;
};

struct Decoder {
    struct Part {
        std::vector<uint32_t> words;   // own aligned copy of the sub-stream
        Reader r;
    };
    std::vector<Part> parts;
    bool has_stream = false;
};

// rans.cpp:279-331
static int decode_part(Reader &r, const int16_t *idx, long long n, const Tables &t, int16_t *out)
{
    for (long long i = 0; i < n; ++i) {
        const int ti = idx[i];
        if (ti < 0 || ti >= t.num) return PMCTF_EINVAL;
        const int32_t *cdf = t.cdf.data() + (size_t)ti * t.stride;
        const int32_t size = t.sizes[ti], escape = size - 2;
        if (escape < 0 || size > t.stride) return PMCTF_EINVAL;
        const uint32_t target = r.peek(PROB_BITS);
        // first entry above the target, minus one.  The reference scans linearly from entry 0 (rans.cpp:296-300); the tables are
        // non-decreasing, so the same entry is found by bisection in 7 probes instead of ~50 (the mode of a table sits in its middle)
        int32_t s = 0, hi = size - 1;
        if (t.monotone) {
            while (s < hi) {
                const int32_t mid = (s + hi + 1) >> 1;
                if ((uint32_t)cdf[mid] <= target) s = mid; else hi = mid - 1;
            }
        } else {
            while (s + 1 < size && (uint32_t)cdf[s + 1] <= target) ++s;
        }
        if (s + 1 >= size) return PMCTF_EINVAL;
        r.advance((uint32_t)cdf[s], (uint32_t)(cdf[s + 1] - cdf[s]), PROB_BITS);
        int32_t v = s;
        if (v == escape) {
            uint32_t d = r.get_raw(RAW_BITS);
            int32_t digits = (int32_t)d;
            while (d == RAW_MAX) {
                d = r.get_raw(RAW_BITS);
                digits += (int32_t)d;
                if (digits > 64) return PMCTF_EINVAL;   // a corrupt stream, not a symbol
            }
            uint32_t raw = 0;
            for (int j = 0; j < digits; ++j) raw |= r.get_raw(RAW_BITS) << (j * RAW_BITS);
            v = (int32_t)(raw >> 1);
            v = (raw & 1u) ? -v - 1 : v + escape;
        }
        out[i] = (int16_t)(v + t.offsets[ti]);
        if (r.overrun) return PMCTF_EINVAL;
    }
    return 0;
}

static int make_tables(Tables &t, const int *cdfs, int num, int stride, const int *sizes, const int *offsets)
{
    if (!cdfs || !sizes || !offsets || num <= 0 || stride < 2) return PMCTF_EINVAL;
    t.num = num;
    t.stride = stride;
    t.cdf.assign(cdfs, cdfs + (size_t)num * stride);
    t.sizes.assign(sizes, sizes + num);
    t.offsets.assign(offsets, offsets + num);
    for (int i = 0; i < num; ++i) {
        if (t.sizes[i] < 2 || t.sizes[i] > stride) return PMCTF_EINVAL;
        const int32_t *row = t.cdf.data() + (size_t)i * stride;
        for (int j = 1; j < t.sizes[i]; ++j)
            if ((uint32_t)row[j] < (uint32_t)row[j - 1]) t.monotone = false;
    }
    return 0;
}

// entropy_models.py:37-40 + 266-270 in one pass: the symbols as int16 (clamped to +-30000, truncated like .to(int16)) and
// the index of the scale in the logarithmic table, idx = int(clamp((log(max(scale, 1e-5)) - log_min) / log_step, 0, levels-1))
__global__ void __launch_bounds__(256) gaussian_symbolize_kernel(const float *__restrict__ sym, const float *__restrict__ scale,
                                                                 long long n, float log_min, float log_step, float top,
                                                                 short *__restrict__ sym16, short *__restrict__ idx16)
{
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 sc = __ldg(reinterpret_cast<const float4 *>(scale) + i);
        const float s[4] = {sc.x, sc.y, sc.z, sc.w};
        short id[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v = (logf(fmaxf(s[k], 1e-5f)) - log_min) / log_step;
            v = fminf(fmaxf(v, 0.0f), top);
            id[k] = (short)(int)v;
        }
        reinterpret_cast<short4 *>(idx16)[i] = make_short4(id[0], id[1], id[2], id[3]);
        if (sym) {
            const float4 y = __ldg(reinterpret_cast<const float4 *>(sym) + i);
            reinterpret_cast<short4 *>(sym16)[i] =
                make_short4((short)(int)fminf(fmaxf(y.x, -30000.0f), 30000.0f), (short)(int)fminf(fmaxf(y.y, -30000.0f), 30000.0f),
                            (short)(int)fminf(fmaxf(y.z, -30000.0f), 30000.0f), (short)(int)fminf(fmaxf(y.w, -30000.0f), 30000.0f));
        }
    }
    if (blockIdx.x == 0) {   // tail
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
            float v = (logf(fmaxf(__ldg(scale + i), 1e-5f)) - log_min) / log_step;
            v = fminf(fmaxf(v, 0.0f), top);
            idx16[i] = (short)(int)v;
            if (sym) sym16[i] = (short)(int)fminf(fmaxf(__ldg(sym + i), -30000.0f), 30000.0f);
        }
    }
}

} // namespace rans
} // namespace pmctf

using namespace pmctf;
using namespace pmctf::rans;

extern "C" {

int pmctf_pmf_to_quantized_cdf(const float *pmf, int n, int precision, unsigned *cdf)
{   // ops.cpp:24-82: frequencies rounded at `precision` bits, renormalised to the exact total, zero-width symbols widened by
    // taking one count from the narrowest symbol that can spare it
    if (!pmf || !cdf || n <= 0 || precision < 1 || precision > 16) return PMCTF_EINVAL;
    const uint32_t one = 1u << precision;
    cdf[0] = 0;
    for (int i = 0; i < n; ++i) cdf[i + 1] = (uint32_t)(roundf(pmf[i] * (float)one) + 0.5);
    uint32_t total = 0;
    for (int i = 0; i <= n; ++i) total += cdf[i];
    if (total == 0) return PMCTF_EINVAL;
    for (int i = 0; i <= n; ++i) cdf[i] = (uint32_t)(((uint64_t)one * cdf[i]) / total);
    for (int i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
    cdf[n] = one;
    for (int i = 0; i < n; ++i) {
        if (cdf[i] != cdf[i + 1]) continue;
        uint32_t best = ~0u;
        int donor = -1;
        for (int j = 0; j < n; ++j) {
            const uint32_t f = cdf[j + 1] - cdf[j];
            if (f > 1 && f < best) {
                best = f;
                donor = j;
            }
        }
        if (donor < 0) return PMCTF_EINVAL;
        if (donor < i)
            for (int j = donor + 1; j <= i; ++j) cdf[j]--;
        else
            for (int j = i + 1; j <= donor; ++j) cdf[j]++;
    }
    return 0;
}

int pmctf_rans_encoder_create(int multi_thread, int stream_part, void **enc)
{
    if (!enc || stream_part < 1 || stream_part > 16) return PMCTF_EINVAL;   // the stream header keeps (parts - 1) in 4 bits
    Encoder *e = new (std::nothrow) Encoder;
    if (!e) return PMCTF_EINVAL;
    const bool threaded = multi_thread != 0 || stream_part > 1;   // py_rans.cpp:13
    for (int i = 0; i < stream_part; ++i) e->parts.emplace_back(new SubEncoder(threaded));
    *enc = e;
    return 0;
}

int pmctf_rans_encoder_destroy(void *enc)
{
    delete static_cast<Encoder *>(enc);
    return 0;
}

int pmctf_rans_encoder_reset(void *enc)
{
    if (!enc) return PMCTF_EINVAL;
    for (auto &p : static_cast<Encoder *>(enc)->parts) {
        SubEncoder *s = p.get();
        s->submit([s] {
            s->queue.clear();
            s->error = 0;
        });
    }
    return 0;
}

int pmctf_rans_encode_with_indexes(void *enc, const short *symbols, const short *indexes, long long n, const int *cdfs, int cdf_num,
                                   int cdf_stride, const int *cdfs_sizes, const int *offsets)
{
    if (!enc || n < 0 || (n > 0 && (!symbols || !indexes))) return PMCTF_EINVAL;
    auto t = std::make_shared<Tables>();
    int e = make_tables(*t, cdfs, cdf_num, cdf_stride, cdfs_sizes, offsets);
    if (e) return e;
    Encoder *E = static_cast<Encoder *>(enc);
    const long long parts = (long long)E->parts.size(), each = n / parts;
    for (long long i = 0; i < parts; ++i) {   // py_rans.cpp:45-59: equal shares, the last part takes the remainder
        const long long cnt = i < parts - 1 ? each : n - each * (parts - 1);
        auto s16 = std::make_shared<std::vector<int16_t>>(symbols + i * each, symbols + i * each + cnt);
        auto i16 = std::make_shared<std::vector<int16_t>>(indexes + i * each, indexes + i * each + cnt);
        SubEncoder *s = E->parts[(size_t)i].get();
        s->submit([s, s16, i16, t] {
            const int rc = queue_symbols(s->queue, s16->data(), i16->data(), (long long)s16->size(), *t);
            if (rc && !s->error) s->error = rc;
        });
    }
    return 0;
}

int pmctf_rans_encode_chunked(void *enc, const short *symbols, const short *indexes, long long n, long long chunk, const int *cdfs,
                              int cdf_num, int cdf_stride, const int *cdfs_sizes, const int *offsets)
{   // the same as ceil(n / chunk) consecutive pmctf_rans_encode_with_indexes calls of `chunk` symbols each (the last one shorter):
    // every chunk is shared out over the sub-streams on its own, which is what a decoder that asks for `chunk` symbols per call sees
    if (!enc || n < 0 || chunk <= 0 || (n > 0 && (!symbols || !indexes))) return PMCTF_EINVAL;
    auto t = std::make_shared<Tables>();
    int e = make_tables(*t, cdfs, cdf_num, cdf_stride, cdfs_sizes, offsets);
    if (e) return e;
    Encoder *E = static_cast<Encoder *>(enc);
    const long long parts = (long long)E->parts.size();
    std::vector<std::shared_ptr<std::vector<int16_t>>> s16((size_t)parts), i16((size_t)parts);
    for (long long i = 0; i < parts; ++i) {
        s16[(size_t)i] = std::make_shared<std::vector<int16_t>>();
        i16[(size_t)i] = std::make_shared<std::vector<int16_t>>();
    }
    for (long long c0 = 0; c0 < n; c0 += chunk) {
        const long long m = std::min(chunk, n - c0), each = m / parts;
        for (long long i = 0; i < parts; ++i) {
            const long long cnt = i < parts - 1 ? each : m - each * (parts - 1);
            s16[(size_t)i]->insert(s16[(size_t)i]->end(), symbols + c0 + i * each, symbols + c0 + i * each + cnt);
            i16[(size_t)i]->insert(i16[(size_t)i]->end(), indexes + c0 + i * each, indexes + c0 + i * each + cnt);
        }
    }
    for (long long i = 0; i < parts; ++i) {
        SubEncoder *sp = E->parts[(size_t)i].get();
        auto a = s16[(size_t)i], b = i16[(size_t)i];
        sp->submit([sp, a, b, t] {
            const int rc = queue_symbols(sp->queue, a->data(), b->data(), (long long)a->size(), *t);
            if (rc && !sp->error) sp->error = rc;
        });
    }
    return 0;
}

int pmctf_rans_encoder_flush(void *enc)
{
    if (!enc) return PMCTF_EINVAL;
    for (auto &p : static_cast<Encoder *>(enc)->parts) {
        SubEncoder *s = p.get();
        s->submit([s] { code_queue(s->queue, s->bytes, s->words); });
    }
    return 0;
}

/* size of flag byte + per-stream sizes + streams (py_rans.cpp:67-113); waits for the queued work */
long long pmctf_rans_encoded_size(void *enc)
{
    if (!enc) return PMCTF_EINVAL;
    Encoder *E = static_cast<Encoder *>(enc);
    size_t total = 1, widest = 0;
    for (size_t i = 0; i < E->parts.size(); ++i) {
        E->parts[i]->drain();
        if (E->parts[i]->error) return E->parts[i]->error;
        total += E->parts[i]->bytes.size();
        if (i + 1 < E->parts.size() && E->parts[i]->bytes.size() > widest) widest = E->parts[i]->bytes.size();
    }
    total += (E->parts.size() - 1) * (widest > 65535 ? 4 : 2);
    return (long long)total;
}

int pmctf_rans_get_encoded_stream(void *enc, unsigned char *out, long long capacity)
{
    const long long need = pmctf_rans_encoded_size(enc);
    if (need < 0) return (int)need;
    if (!out || capacity < need) return PMCTF_EWORKSPACE;
    Encoder *E = static_cast<Encoder *>(enc);
    const size_t parts = E->parts.size();
    size_t widest = 0;
    for (size_t i = 0; i + 1 < parts; ++i) widest = widest > E->parts[i]->bytes.size() ? widest : E->parts[i]->bytes.size();
    const int field = widest > 65535 ? 4 : 2;
    out[0] = (unsigned char)(((parts - 1) << 4) + (field == 2 ? 1 : 0));
    size_t pos = 1;
    for (size_t i = 0; i + 1 < parts; ++i) {
        if (field == 2) {
            const uint16_t v = (uint16_t)E->parts[i]->bytes.size();
            memcpy(out + pos, &v, 2);
        } else {
            const uint32_t v = (uint32_t)E->parts[i]->bytes.size();
            memcpy(out + pos, &v, 4);
        }
        pos += field;
    }
    for (size_t i = 0; i < parts; ++i) {
        memcpy(out + pos, E->parts[i]->bytes.data(), E->parts[i]->bytes.size());
        pos += E->parts[i]->bytes.size();
    }
    return 0;
}

int pmctf_rans_decoder_create(int stream_part, void **dec)
{
    if (!dec || stream_part < 1 || stream_part > 16) return PMCTF_EINVAL;
    Decoder *d = new (std::nothrow) Decoder;
    if (!d) return PMCTF_EINVAL;
    d->parts.resize((size_t)stream_part);
    *dec = d;
    return 0;
}

int pmctf_rans_decoder_destroy(void *dec)
{
    delete static_cast<Decoder *>(dec);
    return 0;
}

int pmctf_rans_decoder_set_stream(void *dec, const unsigned char *bytes, long long n)
{   // py_rans.cpp:129-163
    if (!dec || !bytes || n < 1) return PMCTF_EINVAL;
    Decoder *D = static_cast<Decoder *>(dec);
    const size_t parts = (size_t)(bytes[0] >> 4) + 1;
    if (parts != D->parts.size()) return PMCTF_EINVAL;
    const int field = (bytes[0] & 0x0f) == 1 ? 2 : 4;
    size_t pos = 1, sum = 0;
    std::vector<size_t> sizes;
    for (size_t i = 0; i + 1 < parts; ++i) {
        if (pos + field > (size_t)n) return PMCTF_EINVAL;
        size_t v;
        if (field == 2) {
            uint16_t t;
            memcpy(&t, bytes + pos, 2);
            v = t;
        } else {
            uint32_t t;
            memcpy(&t, bytes + pos, 4);
            v = t;
        }
        pos += field;
        sizes.push_back(v);
        sum += v;
    }
    if (pos + sum > (size_t)n) return PMCTF_EINVAL;
    sizes.push_back((size_t)n - pos - sum);
    for (size_t i = 0; i < parts; ++i) {
        if (sizes[i] < 8 || (sizes[i] & 3)) return PMCTF_EINVAL;   // a flushed stream holds at least the 64-bit final state
        D->parts[i].words.resize(sizes[i] / 4);
        memcpy(D->parts[i].words.data(), bytes + pos, sizes[i]);
        D->parts[i].r.start(D->parts[i].words.data(), D->parts[i].words.size());
        pos += sizes[i];
    }
    D->has_stream = true;
    return 0;
}

int pmctf_rans_decode_stream(void *dec, const short *indexes, long long n, const int *cdfs, int cdf_num, int cdf_stride,
                             const int *cdfs_sizes, const int *offsets, short *out)
{   // py_rans.cpp:165-225: the sub-streams decode concurrently, each continuing from where its last call stopped
    if (!dec || n < 0 || (n > 0 && (!indexes || !out))) return PMCTF_EINVAL;
    Decoder *D = static_cast<Decoder *>(dec);
    if (!D->has_stream) return PMCTF_EINVAL;
    Tables t;
    int e = make_tables(t, cdfs, cdf_num, cdf_stride, cdfs_sizes, offsets);
    if (e) return e;
    const long long parts = (long long)D->parts.size(), each = n / parts;
    std::vector<int> rc((size_t)parts, 0);
    std::vector<std::thread> th;
    for (long long i = 0; i < parts; ++i) {
        const long long cnt = i < parts - 1 ? each : n - each * (parts - 1);
        auto job = [&, i, cnt] { rc[(size_t)i] = decode_part(D->parts[(size_t)i].r, indexes + i * each, cnt, t, out + i * each); };
        if (i + 1 < parts)
            th.emplace_back(job);
        else
            job();
    }
    for (auto &x : th) x.join();
    for (int r : rc)
        if (r) return r;
    return 0;
}

int pmctf_rans_decoder_parts(void *dec)
{
    return dec ? (int)static_cast<Decoder *>(dec)->parts.size() : 0;
}

int pmctf_rans_decoder_peek(void *dec, int part, unsigned long long *x, long long *pos, long long *nwords, const unsigned int **words)
{
    if (!dec || !x || !pos || !nwords || !words) return PMCTF_EINVAL;
    Decoder *D = static_cast<Decoder *>(dec);
    if (!D->has_stream || part < 0 || part >= (int)D->parts.size()) return PMCTF_EINVAL;
    Decoder::Part &p = D->parts[(size_t)part];
    *x = p.r.x;
    *pos = (long long)(p.r.p - p.words.data());
    *nwords = (long long)p.words.size();
    *words = p.words.data();
    return 0;
}

int pmctf_rans_decoder_seek(void *dec, int part, unsigned long long x, long long pos)
{
    if (!dec) return PMCTF_EINVAL;
    Decoder *D = static_cast<Decoder *>(dec);
    if (!D->has_stream || part < 0 || part >= (int)D->parts.size()) return PMCTF_EINVAL;
    Decoder::Part &p = D->parts[(size_t)part];
    if (pos < 2 || pos > (long long)p.words.size()) return PMCTF_EINVAL;
    p.r.x = x;
    p.r.p = p.words.data() + pos;
    return 0;
}

int pmctf_gaussian_symbolize(const float *symbols, const float *scales, long long n, float log_scale_min, float log_scale_step,
                             int scale_levels, short *sym16, short *idx16, void *stream)
{
    if (n == 0) return 0;
    if (!scales || !idx16 || n < 0 || scale_levels < 1 || scale_levels > 32767 || !(log_scale_step > 0.0f)) return PMCTF_EINVAL;
    if ((symbols != nullptr) != (sym16 != nullptr)) return PMCTF_EINVAL;
    if ((((uintptr_t)symbols | (uintptr_t)scales) & 15) || (((uintptr_t)sym16 | (uintptr_t)idx16) & 7)) return PMCTF_EINVAL;
    long long blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    gaussian_symbolize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(symbols, scales, n, log_scale_min, log_scale_step,
                                                                                 (float)(scale_levels - 1), sym16, idx16);
    count_launch();
    return (int)cudaGetLastError();
}

} // extern "C"
