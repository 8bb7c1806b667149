#include "textio.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace dnaldpc {

static bool slurp(const std::string &path, std::string &buf, std::string &err) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { err = "Can't open input file: " + path; return false; }
    char tmp[1 << 16];
    size_t n;
    while ((n = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.append(tmp, n);
    fclose(f);
    return true;
}

bool read_codeword_txt(const std::string &path, int N, std::vector<signed char> &out, std::string &err) {
    std::string buf;
    if (!slurp(path, buf, err)) return false;
    out.assign((size_t)N, 0);
    const char *p = buf.c_str();
    for (int i = 0; i < N; i++) {
        char *end;
        long v = strtol(p, &end, 10);
        if (end == p) { err = "File " + path + " holds fewer than " + std::to_string(N) + " codeword bits"; return false; }
        out[(size_t)i] = (signed char)v;
        p = end;
    }
    return true;
}

bool read_llr_txt(const std::string &path, int N, std::vector<double> &out, std::string &err) {
    std::string buf;
    if (!slurp(path, buf, err)) return false;
    out.assign((size_t)N, 0.0);
    const char *p = buf.c_str();
    for (int i = 0; i < N; i++) {
        char *end;
        double v = strtod(p, &end);  // same decimal->binary64 conversion as fscanf("%lf") (correctly rounded in glibc)
        if (end == p) { err = "File " + path + " holds fewer than " + std::to_string(N) + " LLR values"; return false; }
        out[(size_t)i] = v;
        p = end;
    }
    return true;
}

bool write_dec_txt(const std::string &path, const unsigned char *bits, int N, std::string &err) {
    std::string s;
    s.reserve((size_t)N * 2);
    for (int i = 0; i < N; i++) { s.push_back(bits[i] ? '1' : '0'); s.push_back(' '); }
    FILE *f = fopen(path.c_str(), "w");
    if (!f) { err = "Can't create " + path; return false; }
    bool ok = fwrite(s.data(), 1, s.size(), f) == s.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) err = "Error writing " + path;
    return ok;
}

}  // namespace dnaldpc
