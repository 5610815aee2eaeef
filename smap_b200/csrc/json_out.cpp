// Result serialisation: skeleton records -> the reference's result JSON, byte for byte what
// json.dump(result, f) writes for the dict built by generate_3d_point_pairs / save_result
// (exps/stage3_root2/test.py:32-34,147-152, exps/stage3_root2/test_util.py:146-158).  Host code only.
//
// Python's json module prints floats with float.__repr__ (shortest round-trip digits, exponent form when the decimal
// point position is <= -4 or > 16, "e-05"-style exponents); std::to_chars supplies the shortest digits and the layout
// rule of CPython's format_float_short ('r') is applied on top.  pred_2d values are float32 in the reference
// (ndarray.tolist() widens them to double), so they are widened here before printing.
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/smap_b200.h"

struct smapb_json_writer {
    FILE* f = nullptr;
    std::string buf;
    bool first_pair = true;
    std::string err;
};

namespace {

void put_float(std::string& out, double x) {
    if (std::isnan(x)) { out += "NaN"; return; }
    if (std::isinf(x)) { out += x < 0 ? "-Infinity" : "Infinity"; return; }
    if (x == 0.0) { out += std::signbit(x) ? "-0.0" : "0.0"; return; }
    char tmp[64];
    auto r = std::to_chars(tmp, tmp + sizeof tmp, x, std::chars_format::scientific);  // [-]d[.ddd]e[+-]XX, shortest
    const char* p = tmp;
    if (*p == '-') { out += '-'; p++; }
    char digits[32];
    int nd = 0;
    while (p < r.ptr && *p != 'e') {
        if (*p != '.') digits[nd++] = *p;
        p++;
    }
    p++;  // 'e'
    int e10 = 0;
    const bool eneg = *p == '-';
    p++;
    while (p < r.ptr) e10 = e10 * 10 + (*p++ - '0');
    if (eneg) e10 = -e10;
    const int decpt = e10 + 1;  // value = 0.d1d2... * 10^decpt
    if (decpt <= -4 || decpt > 16) {  // exponent form: d[.ddd]e[+-]XX (at least two exponent digits)
        out += digits[0];
        if (nd > 1) { out += '.'; out.append(digits + 1, nd - 1); }
        char eb[16];
        snprintf(eb, sizeof eb, "e%c%02d", e10 < 0 ? '-' : '+', e10 < 0 ? -e10 : e10);
        out += eb;
    } else if (decpt <= 0) {
        out += "0.";
        out.append((size_t)(-decpt), '0');
        out.append(digits, nd);
    } else if (decpt >= nd) {
        out.append(digits, nd);
        out.append((size_t)(decpt - nd), '0');
        out += ".0";
    } else {
        out.append(digits, decpt);
        out += '.';
        out.append(digits + decpt, nd - decpt);
    }
}

// json.dumps(str) with ensure_ascii=True; input is UTF-8 (invalid sequences are passed through as U+FFFD)
void put_string(std::string& out, const char* s) {
    out += '"';
    const unsigned char* p = reinterpret_cast<const unsigned char*>(s ? s : "");
    char eb[16];
    while (*p) {
        unsigned c = *p;
        if (c == '"') { out += "\\\""; p++; }
        else if (c == '\\') { out += "\\\\"; p++; }
        else if (c == '\n') { out += "\\n"; p++; }
        else if (c == '\r') { out += "\\r"; p++; }
        else if (c == '\t') { out += "\\t"; p++; }
        else if (c == '\b') { out += "\\b"; p++; }
        else if (c == '\f') { out += "\\f"; p++; }
        else if (c < 0x20) { snprintf(eb, sizeof eb, "\\u%04x", c); out += eb; p++; }
        else if (c < 0x80) { out += (char)c; p++; }
        else {
            unsigned cp = 0xFFFD;
            int n = (c >= 0xF0 && c < 0xF8) ? 3 : (c >= 0xE0) ? 2 : (c >= 0xC0) ? 1 : -1;
            if (n > 0) {
                cp = c & (0x3F >> n);
                int k = 1;
                for (; k <= n; k++) {
                    if ((p[k] & 0xC0) != 0x80) break;
                    cp = (cp << 6) | (p[k] & 0x3F);
                }
                if (k <= n) { cp = 0xFFFD; n = 0; }
            } else {
                n = 0;
            }
            p += n + 1;
            if (cp >= 0x10000) {
                cp -= 0x10000;
                snprintf(eb, sizeof eb, "\\u%04x\\u%04x", 0xD800 + (cp >> 10), 0xDC00 + (cp & 0x3FF));
            } else {
                snprintf(eb, sizeof eb, "\\u%04x", cp);
            }
            out += eb;
        }
    }
    out += '"';
}

template <typename T>
void put_bodies(std::string& out, const T (*bodies)[SMAPB_NJ][4], int n) {  // [[[x, y, z, s], ...15], ...n]
    out += '[';
    for (int p = 0; p < n; p++) {
        out += p ? ", [" : "[";
        for (int j = 0; j < SMAPB_NJ; j++) {
            out += j ? ", [" : "[";
            for (int c = 0; c < 4; c++) {
                if (c) out += ", ";
                put_float(out, (double)bodies[p][j][c]);
            }
            out += ']';
        }
        out += ']';
    }
    out += ']';
}

int flush(smapb_json_writer* w) {
    if (!w->buf.empty()) {
        if (fwrite(w->buf.data(), 1, w->buf.size(), w->f) != w->buf.size()) {
            w->err = "write failed";
            return -11;
        }
        w->buf.clear();
    }
    return 0;
}

}  // namespace

extern "C" {

__attribute__((visibility("default"))) int smapb_json_open(smapb_json_writer** out, const char* path, const char* model_pattern) {
    if (!out || !path) return -1;
    FILE* f = fopen(path, "wb");
    if (!f) return -11;
    smapb_json_writer* w = new smapb_json_writer;
    w->f = f;
    w->buf.reserve(1 << 20);
    w->buf += "{\"model_pattern\": ";  // result['model_pattern'] = cfg.DATASET.NAME (test.py:33)
    put_string(w->buf, model_pattern);
    w->buf += ", \"3d_pairs\": [";
    *out = w;
    return 0;
}

// one 'pair' per image with at least one person (test.py:130-131 skips the others), keys in save_result's insertion order
__attribute__((visibility("default"))) int smapb_json_append(smapb_json_writer* w, const smapb_record* rec, int B,
                                                             const char* const* image_paths) {
    if (!w || !w->f || (B > 0 && (!rec || !image_paths))) return -1;
    for (int b = 0; b < B; b++) {
        const int n = rec[b].count < 0 ? 0 : (rec[b].count > SMAPB_MAXP ? SMAPB_MAXP : rec[b].count);
        if (n == 0) continue;
        std::string& o = w->buf;
        o += w->first_pair ? "{" : ", {";
        w->first_pair = false;
        o += "\"pred_2d\": ";
        put_bodies(o, rec[b].pred2d, n);
        o += ", \"pred_3d\": ";
        put_bodies(o, rec[b].pred3d, n);
        o += ", \"root_d\": [";
        for (int p = 0; p < n; p++) {
            if (p) o += ", ";
            put_float(o, rec[b].root_depth[p]);
        }
        o += "], \"image_path\": ";
        put_string(o, image_paths[b]);
        o += ", \"gt_3d\": [], \"gt_2d\": []}";
        if (o.size() > (1u << 20) - 65536) {
            int rc = flush(w);
            if (rc) return rc;
        }
    }
    return 0;
}

__attribute__((visibility("default"))) int smapb_json_close(smapb_json_writer* w) {
    if (!w) return -1;
    int rc = 0;
    if (w->f) {
        w->buf += "]}";
        rc = flush(w);
        if (fclose(w->f) != 0 && !rc) rc = -11;
    }
    delete w;
    return rc;
}

}  // extern "C"
