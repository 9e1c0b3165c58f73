// Feeds the inputs written by gen.py to the host layer's readers; built with -fsanitize=address,undefined by run.sh.
// Nothing here may crash or trip a sanitizer: damaged input is either rejected with a message or read as what it says.
#include "../../include/fwhost.h"
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

static std::string slurp(const std::string &p) { std::ifstream f(p, std::ios::binary); std::stringstream s; s << f.rdbuf(); return s.str(); }

int main(int argc, char **argv)
{
    const std::string d = argc > 1 ? argv[1] : ".";
    char err[512];
    long ok = 0, rejected = 0;
    for (const char *suffix : {".vw.fwcache", ".vw.gz.fwcache"})
        for (int i = 0; i < 1500; i++) {
            uint32_t *recs = nullptr, *off = nullptr; uint64_t nw = 0; char *js = nullptr;
            if (fwhost_cache_read((d + "/m" + std::to_string(i) + suffix).c_str(), nullptr, &recs, &nw, &off, &js, err, sizeof(err)) >= 0) {
                ok++; fwhost_free(recs); fwhost_free(off); fwhost_free(js);
            } else rejected++;
        }
    printf("mutated caches (plain + LZ4 frame): read %ld, rejected %ld\n", ok, rejected);
    const std::string vj = slurp(d + "/vwmap.json");
    void *P = fwhost_parser_new(vj.c_str(), err, sizeof(err));
    if (!P) { puts(err); return 1; }
    std::ifstream f(d + "/lines.txt");
    std::string line;
    std::vector<uint32_t> out(1 << 12);
    ok = rejected = 0;
    while (std::getline(f, line)) {
        line.push_back('\n');
        (fwhost_parser_parse_line(P, line.data(), line.size(), out.data(), out.size(), err, sizeof(err)) > 0 ? ok : rejected)++;
    }
    printf("generated lines: parsed %ld, rejected %ld\n", ok, rejected);
    const std::string l = "1 |A a b c d e f g |D x:2 y:3\n";
    printf("record buffer of 8 words -> %d (%s)\n", fwhost_parser_parse_line(P, l.data(), l.size(), out.data(), 8, err, sizeof(err)), err);
    fwhost_parser_free(P);
    ok = rejected = 0;
    for (int i = 0; i < 2000; i++) {
        void *r = fwhost_regressor_open((d + "/r" + std::to_string(i) + ".fw").c_str(), err, sizeof(err));
        if (!r) { rejected++; continue; }
        float buf[64];
        fwhost_regressor_read(r, buf, sizeof(buf));
        fwhost_regressor_close(r);
        ok++;
    }
    printf("mutated regressor files: header accepted %ld, rejected %ld\n", ok, rejected);
    return 0;
}
