// CPU micro-benchmark of the ASEQ loader: parses every *.ASEQ of a directory against a BED panel into
// the packed wire format, like load_counts of the programs but into plain (not pinned) memory so that it runs without a
// GPU.  Build and run (dev container):
//   g++ -O2 -std=c++17 -pthread -I include scripts/parse_bench.cpp -L amplisolve_b200/lib -lamplisolve_b200 \
//       -Wl,-rpath,$PWD/amplisolve_b200/lib -o /tmp/parse_bench && /tmp/parse_bench panel.bed ASEQ_DIR [threads]
#include "../amplisolve_b200/csrc/as_host.cpp"

// parse_bench --noise-table table.txt pieces [dump.txt] [repeat]: the caller program's noise-table parser alone (as_host.cpp
// parse_noise_table + Panel::link), `pieces` = 0 for the automatic cut.  The dump holds, per row, everything the parser
// stored (the thresholds as float bit patterns), for the CPU test that compares it with an independent reading of the table.
static int noise_table_mode(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: parse_bench --noise-table table.txt pieces [dump.txt] [repeat]\n"); return 2; }
    std::string text;
    if (!read_file(argv[2], text)) { fprintf(stderr, "cannot read %s\n", argv[2]); return 1; }
    const int pieces = atoi(argv[3]);
    const int repeat = argc > 5 ? atoi(argv[5]) : 1;
    double best_parse = 1e30, best_link = 1e30;
    Panel keep;
    std::vector<float> thr;
    std::vector<std::string> germ, dummy;
    for (int r = 0; r < repeat; ++r) {
        Panel panel;
        const double t0 = PhaseTimer::now();
        parse_noise_table(text, pieces, panel, thr, germ, dummy);
        const double t1 = PhaseTimer::now();
        panel.link();
        const double t2 = PhaseTimer::now();
        best_parse = std::min(best_parse, t1 - t0);
        best_link = std::min(best_link, t2 - t1);
        if (r + 1 == repeat) keep = std::move(panel);
    }
    const int64_t P = keep.size();
    if (argc > 4 && argv[4][0] != '-') {
        FILE* o = fopen(argv[4], "w");
        if (!o) return 1;
        for (int64_t i = 0; i < P; ++i) {
            fprintf(o, "%d %s %d %s %s %d", keep.slot_chrom[i], keep.chroms[(size_t)keep.slot_chrom[i]].c_str(), keep.slot_pos[i],
                    keep.pos_text[i].c_str(), keep.ref[i].empty() ? "~" : keep.ref[i].c_str(), (int)keep.dup[i]);
            for (int k = 0; k < 8; ++k) { uint32_t u; memcpy(&u, &thr[(size_t)i * 8 + k], 4); fprintf(o, " %08x", u); }
            for (int k = 0; k < 4; ++k) fprintf(o, " %s", germ[(size_t)i * 4 + k].empty() ? "~" : germ[(size_t)i * 4 + k].c_str());
            fprintf(o, " %d %d %d\n", keep.twin_head[i], keep.twin_next[i], keep.lookup(keep.slot_chrom[i], keep.slot_pos[i]));
        }
        fputs("DUMMY\n", o);
        for (const std::string& d : dummy) fwrite(d.data(), 1, d.size(), o);
        fclose(o);
    }
    printf("{\"rows\": %lld, \"positions\": %lld, \"pieces\": %zu, \"parse_s\": %.5f, \"link_s\": %.5f, \"rows_per_s\": %.3g}\n", (long long)P,
           (long long)keep.n_positions, dummy.size(), best_parse, best_link, P / (best_parse + best_link));
    return 0;
}

// parse_bench --percent-f first_bits last_bits stride: append_percent_f (the noise-table writer's "%f") against snprintf for
// the floats whose bit patterns are first, first + stride, ... <= last.
static int percent_f_mode(int argc, char** argv) {
    if (argc < 5) return 2;
    const uint64_t first = strtoull(argv[2], nullptr, 0), last = strtoull(argv[3], nullptr, 0), stride = std::max(1ull, strtoull(argv[4], nullptr, 0));
    long long tried = 0, bad = 0;
    std::string o;
    char cell[64];
    for (uint64_t u = first; u <= last; u += stride) {
        const uint32_t w = (uint32_t)u;
        float v;
        memcpy(&v, &w, 4);
        o.clear();
        append_percent_f(o, v);
        snprintf(cell, sizeof cell, "%f", (double)v);
        ++tried;
        if (o != cell) ++bad;
    }
    printf("{\"tried\": %lld, \"differences\": %lld}\n", tried, bad);
    return bad ? 1 : 0;
}

// parse_bench --row-scan n seed: n random ASEQ rows -- positions up to 2^31, counts of 1..10 digits, chromosome names of
// 1..24 characters, "\r\n" ends, tokens in the four unused columns -- through the vector scanner, the byte-wise scanner and
// the general parser: whenever a faster one accepts a row it must give the fields of the general one and stop at the same byte.
static int row_scan_mode(int argc, char** argv) {
    if (argc < 4) return 2;
    const long n = atol(argv[2]);
    uint64_t x = strtoull(argv[3], nullptr, 0) * 2654435761ull + 88172645463325252ull;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    std::string text(32, '#');  // the scanners may look at the bytes before a row
    text.back() = '\n';
    std::vector<size_t> starts;
    for (long i = 0; i < n; ++i) {
        starts.push_back(text.size());
        const int cl = 1 + (int)(rnd() % ((rnd() & 7) ? 5 : 24));
        for (int k = 0; k < cl; ++k) text.push_back((char)('A' + rnd() % 26));
        const int pd = 1 + (int)(rnd() % 10);
        text.push_back('\t');
        text += std::to_string(rnd() % (uint64_t)std::min(2147483647.0, pow(10.0, pd)));
        if (rnd() & 3) text += "\t.\t.\t.\t.";
        else for (int k = 0; k < 4; ++k) { text.push_back('\t'); for (int c = 1 + (int)(rnd() % 3); c > 0; --c) text.push_back((char)('!' + rnd() % 90)); }
        for (int k = 0; k < 9; ++k) {
            const int d = (rnd() & 15) ? 1 + (int)(rnd() % 5) : 1 + (int)(rnd() % 10);
            text.push_back('\t');
            text += std::to_string(rnd() % (uint64_t)pow(10.0, d));
        }
        if ((rnd() & 15) == 0) text.push_back('\r');
        text.push_back('\n');
    }
    text.append(512, '\n');
    long long avx_taken = 0, fast_taken = 0, bad = 0;
    for (size_t s0 : starts) {
        const char* e = text.data() + text.size();
        const char* pg = text.data() + s0;
        AseqRow g, f, v;
        const int got = parse_row_general(pg, e, g);
        const char* pf = text.data() + s0;
        const bool okf = parse_row_fast(pf, f);
        auto same = [&](const AseqRow& a, const char* pa) {
            bool eq = pa == pg && a.chrom == g.chrom && a.chrom_len == g.chrom_len && a.pos == g.pos;
            for (int k = 0; k < 9; ++k) eq = eq && a.v[k] == g.v[k];
            return eq;
        };
        if (okf) { ++fast_taken; if (got != 1 || !same(f, pf)) ++bad; }
#ifdef AS_ROW_AVX2
        if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2")) {
            const char* pv = text.data() + s0;
            const bool okv = parse_row_avx2(pv, v);
            if (okv) { ++avx_taken; if (got != 1 || !same(v, pv)) ++bad; }
            else if (pv != text.data() + s0) ++bad;
        }
#endif
    }
    printf("{\"rows\": %ld, \"vector_scanner_took\": %lld, \"bytewise_scanner_took\": %lld, \"differences\": %lld}\n", n, avx_taken, fast_taken, bad);
    return bad ? 1 : 0;
}

// parse_bench --percent-g first_bits last_bits stride: append_percent_g (the noise-table writer's Germ_Max cells) against snprintf
// for the floats whose bit patterns are first, first + stride, ... <= last; values it declines count as `declined`.
static int percent_g_mode(int argc, char** argv) {
    if (argc < 5) return 2;
    const uint64_t first = strtoull(argv[2], nullptr, 0), last = strtoull(argv[3], nullptr, 0), stride = std::max(1ull, strtoull(argv[4], nullptr, 0));
    long long tried = 0, bad = 0, declined = 0;
    std::string o;
    char cell[64];
    for (uint64_t u = first; u <= last; u += stride) {
        const uint32_t w = (uint32_t)u;
        float v;
        memcpy(&v, &w, 4);
        o.clear();
        ++tried;
        if (!append_percent_g(o, v)) { ++declined; if (!o.empty()) ++bad; continue; }
        snprintf(cell, sizeof cell, "%g", (double)v);
        if (o != cell) ++bad;
    }
    printf("{\"tried\": %lld, \"declined\": %lld, \"differences\": %lld}\n", tried, declined, bad);
    return bad ? 1 : 0;
}

// parse_bench --pool-stress rounds: three threads start parallel phases at the same time (run_on_threads: the pool when it is
// free, threads of their own when it is not); every phase must run each of its items exactly once and return.
static int pool_stress_mode(int argc, char** argv) {
    const int rounds = argc > 2 ? atoi(argv[2]) : 5000;
    std::atomic<long long> total{0};
    std::atomic<int> bad{0};
    auto user = [&](int id) {
        for (int r = 0; r < rounds; ++r) {
            const unsigned n = 1 + (unsigned)((r * 7 + id) % 12), items = (unsigned)(r % 50) + 1;
            std::atomic<unsigned> next{0};
            std::atomic<long long> sum{0};
            run_on_threads(n, [&]() {
                for (unsigned k = next.fetch_add(1); k < items; k = next.fetch_add(1)) sum += k + 1;
            });
            if (sum != (long long)items * (items + 1) / 2) ++bad;
            total += sum;
        }
    };
    std::thread a(user, 0), b(user, 1), c(user, 2);
    a.join(); b.join(); c.join();
    printf("{\"phases\": %d, \"items\": %lld, \"bad\": %d}\n", 3 * rounds, (long long)total, (int)bad);
    return bad ? 1 : 0;
}

// parse_bench --stof n seed: threshold_text_to_float (the noise-table parser's std::stof) against strtof: every "d.dddddd"
// with a value below n / 10^6 (what "%f" writes), then n random texts [-]digits[.digits] of up to 10 digits, exponents and
// words that must go to strtof.
static int stof_mode(int argc, char** argv) {
    if (argc < 4) return 2;
    const long n = atol(argv[2]);
    uint64_t x = strtoull(argv[3], nullptr, 0) * 2654435761ull + 88172645463325252ull;
    auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
    long long tried = 0, bad = 0;
    char buf[64];
    auto check = [&](const char* t, size_t len) {
        const float a = threshold_text_to_float(t, t + len), b = len ? strtof(t, nullptr) : 0.f;
        ++tried;
        if (memcmp(&a, &b, 4) != 0) ++bad;
    };
    for (long k = 0; k < n; ++k) check(buf, (size_t)snprintf(buf, sizeof buf, "%ld.%06ld", k / 1000000, k % 1000000));
    for (long k = 0; k < n; ++k) {
        const int digits = 1 + (int)(rnd() % 10), frac = (int)(rnd() % (unsigned)(digits + 1));
        std::string d = std::to_string(rnd() % (uint64_t)pow(10.0, digits));
        d.insert(0, (size_t)digits - d.size(), '0');
        std::string t = (rnd() & 1) ? "-" : "";
        t.append(d, 0, (size_t)(digits - frac));
        if (frac || (rnd() & 1)) t += ".";
        t.append(d, (size_t)(digits - frac), std::string::npos);
        if ((rnd() & 31) == 0) t += "e-0" + std::to_string(rnd() % 9);
        check(t.c_str(), t.size());
    }
    for (const char* w : {"nan", "inf", "-inf", "", "-", "+3", ".", "1e5", "0x10", "-0", "-0.000000"}) check(w, strlen(w));
    printf("{\"tried\": %lld, \"differences\": %lld}\n", tried, bad);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && strcmp(argv[1], "--stof") == 0) return stof_mode(argc, argv);
    if (argc >= 2 && strcmp(argv[1], "--pool-stress") == 0) return pool_stress_mode(argc, argv);
    if (argc >= 2 && strcmp(argv[1], "--percent-g") == 0) return percent_g_mode(argc, argv);
    if (argc >= 2 && strcmp(argv[1], "--row-scan") == 0) return row_scan_mode(argc, argv);
    if (argc >= 2 && strcmp(argv[1], "--noise-table") == 0) return noise_table_mode(argc, argv);
    if (argc >= 2 && strcmp(argv[1], "--percent-f") == 0) return percent_f_mode(argc, argv);
    if (argc < 3) { fprintf(stderr, "usage: parse_bench panel.bed aseq_dir [repeat]\n"); return 2; }
    Panel panel;
    int n_amp = 0;
    if (!load_bed(argv[1], panel, n_amp)) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    panel.link();
    std::vector<CountFile> files;
    std::vector<std::string> listed;
    if (!list_count_files(argv[2], files, listed)) { fprintf(stderr, "cannot list %s\n", argv[2]); return 1; }
    const int repeat = argc > 3 ? atoi(argv[3]) : 3;
    const size_t words = files.size() * 2 * (size_t)panel.size();
    uint32_t* mem = (uint32_t*)malloc(std::max<size_t>(16, words * 4));
    std::vector<AseqStats> stats;
    double best = 1e30;
    int64_t rows = 0, bytes = 0;
    for (int r = 0; r < repeat; ++r) {
        const double t0 = PhaseTimer::now();
        if (!load_all<1>(files, 0, files.size(), panel, mem, stats)) return 1;
        best = std::min(best, PhaseTimer::now() - t0);
    }
    int64_t escaped = 0, outside = 0;
    for (const AseqStats& s : stats) { rows += s.rows; escaped += (int64_t)s.wide.size(); outside += s.outside; }
    for (const CountFile& f : files) { struct stat sb; if (stat(f.path.c_str(), &sb) == 0) bytes += sb.st_size; }
    uint64_t sum = 0;
    for (size_t i = 0; i < words; ++i) sum += mem[i];
    printf("{\"files\": %zu, \"slots\": %lld, \"rows\": %lld, \"bytes\": %lld, \"seconds\": %.4f, \"rows_per_s\": %.3g, \"GB_per_s\": %.3f, "
           "\"threads\": %u, \"escaped\": %lld, \"outside\": %lld, \"checksum\": %llu}\n",
           files.size(), (long long)panel.size(), (long long)rows, (long long)bytes, best, rows / best, bytes / best / 1e9,
           std::min<unsigned>(std::thread::hardware_concurrency(), (unsigned)files.size()), (long long)escaped, (long long)outside,
           (unsigned long long)sum);
    free(mem);
    return 0;
}
