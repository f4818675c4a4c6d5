// codec.cpp -- host side of the msbwt RLE byte format and its `.npy` container, the data format either side of
// the query path: what `msbwt2-convert` / `msbwt2-build` write and `load_numpy_file` reads.
//
// Reference: src/bwt_converter.rs -- convert_to_vec (:26-80: a text BWT over `$ACGNT`, newlines ignored, ->
// RLE bytes `symbol | digit << 3`, one byte per base-32 digit of the run length, least significant first),
// save_bwt_numpy (:102-130) and save_bwt_runs_numpy (:152-184): a 96-byte header -- magic, version 1.0, header
// length 0x56, the dict `{'descr': '|u1', 'fortran_order': False, 'shape': (<bytes>, ), }`, padded with spaces and
// ended by a newline -- followed by the RLE bytes.  Plain C++ (no CUDA): marshalling only.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/msbwt_gpu.h"

namespace msbwt {
int codec_fail(int code, const std::string &msg);  // capi.cu: records msbwt_last_error() for the calling thread
}

namespace {

constexpr int kLetterBits = 3, kNumberBits = 5;  // src/msbwt_core.rs:8-14
constexpr uint64_t kDigitMask = (1u << kNumberBits) - 1u;
constexpr size_t kHeaderBytes = 96;

inline int symbol_code(uint8_t ch) {  // `$ACGNT` = 0..5, anything else -1
    switch (ch) {
        case '$': return 0;
        case 'A': return 1;
        case 'C': return 2;
        case 'G': return 3;
        case 'N': return 4;
        case 'T': return 5;
        default: return -1;
    }
}

inline uint64_t digits_of(uint64_t count) {
    uint64_t d = 0;
    for (; count; count >>= kNumberBits) d++;
    return d;
}

inline uint8_t *put_run(uint8_t *at, int sym, uint64_t count) {
    for (; count; count >>= kNumberBits) *at++ = (uint8_t)(sym | ((count & kDigitMask) << kLetterBits));
    return at;
}

// Walks the text once, calling run(sym, count) for every maximal run; newlines neither end nor extend a run.
// Returns the offset of the first unexpected byte, or n.
template <class F>
uint64_t for_each_text_run(const uint8_t *text, uint64_t n, F &&run) {
    int cur = -1;
    uint64_t count = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t ch = text[i];
        if (ch == '\n') continue;
        const int s = symbol_code(ch);
        if (s < 0) return i;
        if (s == cur) {
            count++;
        } else {
            if (count) run(cur, count);
            cur = s;
            count = 1;
        }
    }
    if (count) run(cur, count);
    return n;
}

void fill_header(uint8_t *hdr, uint64_t payload_bytes) {
    memset(hdr, ' ', kHeaderBytes - 1);
    hdr[kHeaderBytes - 1] = '\n';
    static const char lead[] = "\x93NUMPY\x01\x00\x56\x00{'descr': '|u1', 'fortran_order': False, 'shape': (";
    const size_t lead_len = sizeof(lead) - 1;
    memcpy(hdr, lead, lead_len);
    const std::string num = std::to_string(payload_bytes);
    memcpy(hdr + lead_len, num.data(), num.size());
    memcpy(hdr + lead_len + num.size(), ", ), }", 6);
}

int write_npy(const char *path, const uint8_t *payload, uint64_t len) {
    if (!path) return msbwt::codec_fail(MSBWT_EINVAL, "NULL path");
    FILE *f = fopen(path, "wb");
    if (!f) return msbwt::codec_fail(MSBWT_EIO, std::string(path) + ": " + strerror(errno));
    uint8_t hdr[kHeaderBytes];
    fill_header(hdr, len);
    bool ok = fwrite(hdr, 1, kHeaderBytes, f) == kHeaderBytes;
    if (ok && len) ok = fwrite(payload, 1, len, f) == len;
    const int err = errno;
    if (fclose(f) != 0) ok = false;
    if (!ok) return msbwt::codec_fail(MSBWT_EIO, std::string(path) + ": write failed: " + strerror(err ? err : errno));
    return MSBWT_OK;
}

}  // namespace

extern "C" int msbwt_convert_to_rle(const uint8_t *text, uint64_t n, uint8_t **rle, uint64_t *rle_len) {
    if (!rle || !rle_len || (n && !text)) return msbwt::codec_fail(MSBWT_EINVAL, "NULL buffer");
    *rle = nullptr;
    *rle_len = 0;
    uint64_t bytes = 0;
    const uint64_t stop = for_each_text_run(text, n, [&](int, uint64_t count) { bytes += digits_of(count); });
    if (stop != n)  // the reference panics here (src/bwt_converter.rs:43-46)
        return msbwt::codec_fail(MSBWT_EFORMAT, "unexpected symbol (byte " + std::to_string((unsigned)text[stop]) + ") at offset " +
                                                    std::to_string(stop) + " of the text BWT");
    uint8_t *out = (uint8_t *)malloc(bytes ? bytes : 1);
    if (!out) return msbwt::codec_fail(MSBWT_ENOMEM, "RLE output buffer");
    uint8_t *at = out;
    for_each_text_run(text, n, [&](int sym, uint64_t count) { at = put_run(at, sym, count); });
    *rle = out;
    *rle_len = bytes;
    return MSBWT_OK;
}

extern "C" int msbwt_save_rle_npy(const uint8_t *rle, uint64_t len, const char *path) {
    if (len && !rle) return msbwt::codec_fail(MSBWT_EINVAL, "NULL buffer");
    return write_npy(path, rle, len);
}

extern "C" int msbwt_save_runs_npy(const uint8_t *syms, const uint64_t *counts, uint64_t nruns, const char *path) {
    if (nruns && (!syms || !counts)) return msbwt::codec_fail(MSBWT_EINVAL, "NULL buffer");
    uint64_t bytes = 0;
    for (uint64_t i = 0; i < nruns; i++) {
        if (syms[i] >= 6) return msbwt::codec_fail(MSBWT_EINVAL, "run " + std::to_string(i) + " has symbol >= 6");
        bytes += digits_of(counts[i]);
    }
    uint8_t *buf = (uint8_t *)malloc(bytes ? bytes : 1);
    if (!buf) return msbwt::codec_fail(MSBWT_ENOMEM, "RLE output buffer");
    uint8_t *at = buf;
    for (uint64_t i = 0; i < nruns; i++) at = put_run(at, syms[i], counts[i]);
    const int rc = write_npy(path, buf, bytes);
    free(buf);
    return rc;
}
