// jpeg_baseline.hpp — baseline (sequential, Huffman, 8-bit) JPEG decoder: what TextureImage::try_new
// needs for resources/images/earthmap.jpg (src/texture/image.rs:17-25 decodes through the `image`
// crate; that crate is not available to a C++ host and the image has no libjpeg in it either).
// Supports 1 or 3 components, any integer sampling factors (nearest-neighbour chroma upsampling),
// restart intervals; rejects progressive and arithmetic-coded files.  IDCT in double precision,
// JFIF YCbCr -> RGB.  Output is RGBA8, alpha 255 (`into_rgba8`).
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace jpeg_baseline {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Image {
    int width = 0, height = 0;
    std::vector<uint8_t> rgba;
};

namespace detail {

struct Huff {
    uint8_t bits[17] = {0};
    uint8_t vals[256] = {0};
    int mincode[17], maxcode[18], valptr[17];
    bool present = false;
    void build() {
        int code = 0, k = 0;
        for (int l = 1; l <= 16; ++l) {
            valptr[l] = k;
            mincode[l] = code;
            code += bits[l];
            k += bits[l];
            maxcode[l] = bits[l] ? code - 1 : -1;
            code <<= 1;
        }
        maxcode[17] = 0x7fffffff;
        present = true;
    }
};

struct Component {
    int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
    int pred = 0;
    int bw = 0, bh = 0;            // blocks per row / column (padded to MCU multiples)
    std::vector<uint8_t> plane;   // bw*8 x bh*8 samples
};

struct Reader {
    const uint8_t* p;
    size_t n, i = 0;
    uint32_t acc = 0;
    int cnt = 0;
    Reader(const uint8_t* d, size_t len) : p(d), n(len) {}
    int byte() { if (i >= n) throw Error("unexpected end of JPEG data"); return p[i++]; }
    int word() { int a = byte(); return (a << 8) | byte(); }
    int starved = 0;                  // bytes of padding fed because the entropy data ran into a marker or the end
    int bit() {
        if (cnt == 0) {
            bool real = i < n;
            int b = real ? p[i++] : 0;
            if (b == 0xFF) {
                int b2 = i < n ? p[i] : 0;
                if (b2 == 0) ++i;                 // stuffed zero
                else { --i; b = 0; real = false; }   // a marker: feed zeros, leave it for the caller
            }
            // a decoder may look a few bits past the last code; a scan that keeps reading there is truncated
            if (!real && ++starved > 8) throw Error("premature end of JPEG entropy data");
            acc = (uint32_t)b; cnt = 8;
        }
        --cnt;
        return (acc >> cnt) & 1;
    }
    int bitsn(int k) { int v = 0; while (k--) v = (v << 1) | bit(); return v; }
    void reset_bits() { cnt = 0; }
};

inline int decode_symbol(Reader& r, const Huff& h) {
    int code = 0;
    for (int l = 1; l <= 16; ++l) {
        code = (code << 1) | r.bit();
        if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
    throw Error("bad Huffman code");
}

inline int extend(int v, int t) { return v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }

static const int ZIGZAG[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
                               41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
                               15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

inline void idct8x8(const double in[64], uint8_t* out, int stride) {
    static double c[8][8];
    static bool init = false;
    if (!init) {
        for (int x = 0; x < 8; ++x)
            for (int u = 0; u < 8; ++u) c[x][u] = (u == 0 ? std::sqrt(0.125) : 0.5) * std::cos((2 * x + 1) * u * M_PI / 16.0);
        init = true;
    }
    double tmp[64];
    for (int y = 0; y < 8; ++y)        // rows: tmp[y][x] = sum_u c[x][u] in[y][u]
        for (int x = 0; x < 8; ++x) {
            double s = 0;
            for (int u = 0; u < 8; ++u) s += c[x][u] * in[y * 8 + u];
            tmp[y * 8 + x] = s;
        }
    for (int x = 0; x < 8; ++x)
        for (int y = 0; y < 8; ++y) {
            double s = 0;
            for (int v = 0; v < 8; ++v) s += c[y][v] * tmp[v * 8 + x];
            const double px = std::floor(s + 128.5);
            out[y * stride + x] = (uint8_t)(px < 0 ? 0 : (px > 255 ? 255 : px));
        }
}

}  // namespace detail

inline Image decode(const uint8_t* data, size_t len) {
    using namespace detail;
    Reader r(data, len);
    if (r.byte() != 0xFF || r.byte() != 0xD8) throw Error("not a JPEG file");
    uint16_t qt[4][64] = {{0}};
    Huff dc[4], ac[4];
    std::vector<Component> comps;
    int width = 0, height = 0, restart = 0, hmax = 1, vmax = 1;
    for (;;) {
        int b = r.byte();
        if (b != 0xFF) continue;
        int m = r.byte();
        while (m == 0xFF) m = r.byte();
        if (m == 0xD8 || (m >= 0xD0 && m <= 0xD7) || m == 0x01) continue;
        if (m == 0xD9) throw Error("no scan in JPEG file");
        int seglen = r.word();
        size_t end = r.i + seglen - 2;
        if (m == 0xDB) {
            while (r.i < end) {
                int pq_tq = r.byte(), pq = pq_tq >> 4, tq = pq_tq & 15;
                if (tq > 3) throw Error("bad quantisation table id");
                for (int k = 0; k < 64; ++k) qt[tq][ZIGZAG[k]] = (uint16_t)(pq ? r.word() : r.byte());
            }
        } else if (m == 0xC4) {
            while (r.i < end) {
                int tc_th = r.byte(), tc = tc_th >> 4, th = tc_th & 15;
                if (th > 3) throw Error("bad Huffman table id");
                Huff& h = tc ? ac[th] : dc[th];
                int total = 0;
                for (int l = 1; l <= 16; ++l) { h.bits[l] = (uint8_t)r.byte(); total += h.bits[l]; }
                if (total > 256) throw Error("bad Huffman table");
                for (int k = 0; k < total; ++k) h.vals[k] = (uint8_t)r.byte();
                h.build();
            }
        } else if (m == 0xC0 || m == 0xC1) {
            if (r.byte() != 8) throw Error("only 8-bit JPEG is supported");
            height = r.word(); width = r.word();
            if (width <= 0 || height <= 0 || (long long)width * height > (1LL << 28)) throw Error("bad image size");
            int nc = r.byte();
            if (nc != 1 && nc != 3) throw Error("only 1- or 3-component JPEG is supported");
            comps.resize(nc);
            for (auto& c : comps) {
                c.id = r.byte();
                int hv = r.byte();
                c.h = hv >> 4; c.v = hv & 15; c.tq = r.byte();
                if (c.h < 1 || c.v < 1 || c.h > 4 || c.v > 4 || c.tq > 3) throw Error("bad frame component");
                hmax = c.h > hmax ? c.h : hmax; vmax = c.v > vmax ? c.v : vmax;
            }
        } else if (m == 0xC2 || (m >= 0xC5 && m <= 0xCF && m != 0xC8)) {
            throw Error("progressive / arithmetic / lossless JPEG is not supported");
        } else if (m == 0xDD) {
            restart = r.word();
        } else if (m == 0xDA) {
            if (comps.empty() || width <= 0 || height <= 0) throw Error("scan before frame header");
            int ns = r.byte();
            if (ns != (int)comps.size()) throw Error("non-interleaved scans are not supported");
            for (int k = 0; k < ns; ++k) {
                int id = r.byte(), t = r.byte();
                if ((t >> 4) > 3 || (t & 15) > 3) throw Error("bad entropy table selector in the scan header");   // dc[] / ac[] hold 4 tables
                bool found = false;
                for (auto& c : comps) if (c.id == id) { c.td = t >> 4; c.ta = t & 15; found = true; }
                if (!found) throw Error("scan component is not in the frame");
            }
            r.i = end;
            break;
        }
        r.i = end;
    }
    const int mcu_w = 8 * hmax, mcu_h = 8 * vmax;
    const int mcus_x = (width + mcu_w - 1) / mcu_w, mcus_y = (height + mcu_h - 1) / mcu_h;
    for (auto& c : comps) {
        c.bw = mcus_x * c.h; c.bh = mcus_y * c.v;
        c.plane.assign((size_t)c.bw * 8 * c.bh * 8, 0);
        if (!dc[c.td].present || !ac[c.ta].present) throw Error("missing Huffman table");
    }
    int to_restart = restart;
    for (int my = 0; my < mcus_y; ++my)
        for (int mx = 0; mx < mcus_x; ++mx) {
            if (restart && to_restart == 0) {
                r.reset_bits();
                r.starved = 0;
                // skip to the RSTn marker
                while (r.i + 1 < r.n && !(r.p[r.i] == 0xFF && r.p[r.i + 1] >= 0xD0 && r.p[r.i + 1] <= 0xD7)) ++r.i;
                r.i += 2;
                for (auto& c : comps) c.pred = 0;
                to_restart = restart;
            }
            for (auto& c : comps)
                for (int by = 0; by < c.v; ++by)
                    for (int bx = 0; bx < c.h; ++bx) {
                        double coef[64] = {0};
                        int t = decode_symbol(r, dc[c.td]);
                        if (t < 0 || t > 11) throw Error("bad DC coefficient size");   // baseline: at most 11 bits (extend() shifts by it)
                        int diff = t ? extend(r.bitsn(t), t) : 0;
                        c.pred += diff;
                        coef[0] = (double)c.pred * qt[c.tq][0];
                        for (int k = 1; k < 64;) {
                            int rs = decode_symbol(r, ac[c.ta]);
                            int run = rs >> 4, size = rs & 15;
                            if (size == 0) {
                                if (run == 15) { k += 16; continue; }
                                break;
                            }
                            k += run;
                            if (k > 63) throw Error("bad AC coefficient index");
                            coef[ZIGZAG[k]] = (double)extend(r.bitsn(size), size) * qt[c.tq][ZIGZAG[k]];
                            ++k;
                        }
                        const int stride = c.bw * 8;
                        idct8x8(coef, c.plane.data() + (size_t)((my * c.v + by) * 8) * stride + (mx * c.h + bx) * 8, stride);
                    }
            if (restart) --to_restart;
        }
    Image img;
    img.width = width; img.height = height;
    img.rgba.resize((size_t)width * height * 4);
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            double v[3] = {0, 128, 128};
            for (size_t k = 0; k < comps.size(); ++k) {
                const Component& c = comps[k];
                const int sx = x * c.h / hmax, sy = y * c.v / vmax;
                v[k] = c.plane[(size_t)sy * c.bw * 8 + sx];
            }
            double R, G, B;
            if (comps.size() == 1) R = G = B = v[0];
            else {
                R = v[0] + 1.402 * (v[2] - 128.0);
                G = v[0] - 0.344136 * (v[1] - 128.0) - 0.714136 * (v[2] - 128.0);
                B = v[0] + 1.772 * (v[1] - 128.0);
            }
            auto clamp8 = [](double q) { q = std::floor(q + 0.5); return (uint8_t)(q < 0 ? 0 : (q > 255 ? 255 : q)); };
            uint8_t* o = img.rgba.data() + 4 * ((size_t)y * width + x);
            o[0] = clamp8(R); o[1] = clamp8(G); o[2] = clamp8(B); o[3] = 255;
        }
    return img;
}

}  // namespace jpeg_baseline
