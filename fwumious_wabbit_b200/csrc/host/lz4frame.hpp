// lz4frame.hpp -- LZ4 frame format (magic 0x184D2204) reader and writer, written from the published format
// descriptions (lz4_Frame_format.md, lz4_Block_format.md) because the image has no liblz4.
//
// The reference stores the input cache of a `*.gz` input as an LZ4 frame (cache.rs:71, 89-125: `lz4::Decoder` /
// `lz4::EncoderBuilder::new().level(3)` around the same byte stream an uncompressed cache holds).  The decoder accepts
// everything a conforming encoder may produce: linked or independent blocks, block / content checksums, content size,
// dictionary id (rejected: needs an external dictionary), uncompressed blocks, concatenated and skippable frames.
// The encoder writes independent 4 MiB blocks with a greedy hash-chain-free matcher (one 64 K-entry hash table), no
// optional fields; any LZ4 frame decoder -- the reference's included -- reads it.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace fwhost {
namespace lz4 {

inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// xxHash32 (the frame format's header and content checksums)
inline uint32_t xxh32(const uint8_t *p, size_t len, uint32_t seed)
{
    const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
    const uint8_t *end = p + len;
    uint32_t h;
    if (len >= 16) {
        uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        const uint8_t *limit = end - 16;
        do {
            v1 = rotl(v1 + rd32(p) * P2, 13) * P1; p += 4;
            v2 = rotl(v2 + rd32(p) * P2, 13) * P1; p += 4;
            v3 = rotl(v3 + rd32(p) * P2, 13) * P1; p += 4;
            v4 = rotl(v4 + rd32(p) * P2, 13) * P1; p += 4;
        } while (p <= limit);
        h = rotl(v1, 1) + rotl(v2, 7) + rotl(v3, 12) + rotl(v4, 18);
    } else h = seed + P5;
    h += (uint32_t)len;
    while (p + 4 <= end) { h = rotl(h + rd32(p) * P3, 17) * P4; p += 4; }
    while (p < end) { h = rotl(h + (*p) * P5, 11) * P1; p++; }
    h ^= h >> 15; h *= P2; h ^= h >> 13; h *= P3; h ^= h >> 16;
    return h;
}

// one LZ4 block appended to `out`; matches may reach back into everything already in `out` (linked blocks)
inline void decode_block(const uint8_t *src, size_t n, std::vector<uint8_t> &out)
{
    const uint8_t *ip = src, *iend = src + n;
    while (ip < iend) {
        const uint8_t token = *ip++;
        size_t lit = token >> 4;
        if (lit == 15) { uint8_t b; do { if (ip >= iend) throw std::runtime_error("lz4: truncated literal length"); b = *ip++; lit += b; } while (b == 255); }
        if ((size_t)(iend - ip) < lit) throw std::runtime_error("lz4: literals run past the block");
        out.insert(out.end(), ip, ip + lit);
        ip += lit;
        if (ip >= iend) break; // the last sequence is literals only
        if (iend - ip < 2) throw std::runtime_error("lz4: truncated match offset");
        const size_t offset = ip[0] | ((size_t)ip[1] << 8);
        ip += 2;
        size_t mlen = token & 15;
        if (mlen == 15) { uint8_t b; do { if (ip >= iend) throw std::runtime_error("lz4: truncated match length"); b = *ip++; mlen += b; } while (b == 255); }
        mlen += 4;
        if (offset == 0 || offset > out.size()) throw std::runtime_error("lz4: match offset outside the window");
        const size_t at = out.size(), from = at - offset;
        out.resize(at + mlen);
        uint8_t *o = out.data();
        if (offset >= mlen) memcpy(o + at, o + from, mlen);                      // source and destination do not overlap
        else for (size_t i = 0; i < mlen; i++) o[at + i] = o[from + i];          // overlapping match: the pattern repeats
    }
}

inline std::vector<uint8_t> decode_frames(const uint8_t *src, size_t n)
{
    std::vector<uint8_t> out;
    size_t pos = 0;
    bool any = false;
    while (pos < n) {
        if (n - pos < 4) throw std::runtime_error("lz4: truncated frame magic");
        const uint32_t magic = rd32(src + pos);
        pos += 4;
        if ((magic & 0xfffffff0u) == 0x184D2A50u) { // skippable frame
            if (n - pos < 4) throw std::runtime_error("lz4: truncated skippable frame");
            const uint32_t sz = rd32(src + pos);
            pos += 4;
            if (n - pos < sz) throw std::runtime_error("lz4: truncated skippable frame");
            pos += sz;
            continue;
        }
        if (magic != 0x184D2204u) throw std::runtime_error("lz4: not an LZ4 frame");
        any = true;
        if (n - pos < 3) throw std::runtime_error("lz4: truncated frame descriptor");
        const size_t desc = pos;
        const uint8_t flg = src[pos++], bd = src[pos++];
        if ((flg >> 6) != 1) throw std::runtime_error("lz4: unsupported frame version");
        const bool block_checksum = flg & 0x10, content_size = flg & 0x08, content_checksum = flg & 0x04, dict_id = flg & 0x01;
        (void)bd;
        if (content_size) pos += 8;
        if (dict_id) throw std::runtime_error("lz4: frames that need a dictionary are not supported");
        if (pos >= n) throw std::runtime_error("lz4: truncated frame descriptor");
        const uint8_t hc = src[pos];
        if (hc != ((xxh32(src + desc, pos - desc, 0) >> 8) & 0xff)) throw std::runtime_error("lz4: frame descriptor checksum mismatch");
        pos++;
        const size_t frame_start = out.size();
        for (;;) {
            if (n - pos < 4) throw std::runtime_error("lz4: truncated block header");
            const uint32_t bs = rd32(src + pos);
            pos += 4;
            if (bs == 0) break; // EndMark
            const uint32_t len = bs & 0x7fffffffu;
            if (n - pos < len) throw std::runtime_error("lz4: truncated block");
            if (bs & 0x80000000u) out.insert(out.end(), src + pos, src + pos + len); // stored uncompressed
            else decode_block(src + pos, len, out);
            pos += len;
            if (block_checksum) {
                if (n - pos < 4) throw std::runtime_error("lz4: truncated block checksum");
                if (rd32(src + pos) != xxh32(src + pos - len, len, 0)) throw std::runtime_error("lz4: block checksum mismatch");
                pos += 4;
            }
        }
        if (content_checksum) {
            if (n - pos < 4) throw std::runtime_error("lz4: truncated content checksum");
            if (rd32(src + pos) != xxh32(out.data() + frame_start, out.size() - frame_start, 0)) throw std::runtime_error("lz4: content checksum mismatch");
            pos += 4;
        }
    }
    if (!any) throw std::runtime_error("lz4: not an LZ4 frame");
    return out;
}

// greedy LZ4 block compressor (block format rules: the last 5 bytes are literals, no match starts in the last 12 bytes)
inline void encode_block(const uint8_t *src, size_t n, std::vector<uint8_t> &out)
{
    constexpr int HASH_LOG = 16;
    std::vector<uint32_t> table((size_t)1 << HASH_LOG, 0xffffffffu);
    auto hash = [&](uint32_t v) { return (v * 2654435761u) >> (32 - HASH_LOG); };
    auto emit = [&](const uint8_t *lit, size_t lit_len, size_t match_len, size_t offset) { // match_len 0 = final literals
        const size_t ml = match_len ? match_len - 4 : 0;
        out.push_back((uint8_t)((lit_len >= 15 ? 15 : lit_len) << 4 | (match_len ? (ml >= 15 ? 15 : ml) : 0)));
        if (lit_len >= 15) { size_t r = lit_len - 15; while (r >= 255) { out.push_back(255); r -= 255; } out.push_back((uint8_t)r); }
        out.insert(out.end(), lit, lit + lit_len);
        if (match_len) {
            out.push_back((uint8_t)(offset & 0xff)); out.push_back((uint8_t)(offset >> 8));
            if (ml >= 15) { size_t r = ml - 15; while (r >= 255) { out.push_back(255); r -= 255; } out.push_back((uint8_t)r); }
        }
    };
    size_t anchor = 0, i = 0;
    if (n >= 13) {
        const size_t match_limit = n - 12; // no match may start at or after this position
        while (i < match_limit) {
            const uint32_t v = rd32(src + i), h = hash(v);
            const uint32_t cand = table[h];
            table[h] = (uint32_t)i;
            if (cand != 0xffffffffu && i - cand <= 65535 && rd32(src + cand) == v) {
                size_t len = 4;
                const size_t max_len = n - 5 - i; // the last 5 bytes stay literals
                while (len < max_len && src[cand + len] == src[i + len]) len++;
                emit(src + anchor, i - anchor, len, i - cand);
                i += len;
                anchor = i;
            } else i++;
        }
    }
    emit(src + anchor, n - anchor, 0, 0);
}

inline std::vector<uint8_t> encode_frame(const uint8_t *src, size_t n)
{
    std::vector<uint8_t> out;
    const uint8_t hdr[6] = {0x04, 0x22, 0x4D, 0x18, 0x60 /* version 01, independent blocks */, 0x70 /* 4 MiB blocks */};
    out.insert(out.end(), hdr, hdr + 6);
    out.push_back((uint8_t)((xxh32(hdr + 4, 2, 0) >> 8) & 0xff));
    constexpr size_t BLOCK = (size_t)4 << 20;
    std::vector<uint8_t> blk;
    for (size_t pos = 0; pos < n; pos += BLOCK) {
        const size_t len = n - pos < BLOCK ? n - pos : BLOCK;
        blk.clear();
        encode_block(src + pos, len, blk);
        uint32_t word;
        if (blk.size() < len) { word = (uint32_t)blk.size(); out.insert(out.end(), (uint8_t *)&word, (uint8_t *)&word + 4); out.insert(out.end(), blk.begin(), blk.end()); }
        else { word = (uint32_t)len | 0x80000000u; out.insert(out.end(), (uint8_t *)&word, (uint8_t *)&word + 4); out.insert(out.end(), src + pos, src + pos + len); }
    }
    const uint32_t end_mark = 0;
    out.insert(out.end(), (const uint8_t *)&end_mark, (const uint8_t *)&end_mark + 4);
    return out;
}

} // namespace lz4
} // namespace fwhost
