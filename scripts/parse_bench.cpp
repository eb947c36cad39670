// CPU micro-benchmark of the ASEQ loader: parses every *.ASEQ of a directory against a BED panel into
// the packed wire format, like load_counts of the programs but into plain (not pinned) memory so that it runs without a
// GPU.  Build and run (dev container):
//   g++ -O2 -std=c++17 -pthread -I include scripts/parse_bench.cpp -L amplisolve_b200/lib -lamplisolve_b200 \
//       -Wl,-rpath,$PWD/amplisolve_b200/lib -o /tmp/parse_bench && /tmp/parse_bench panel.bed ASEQ_DIR [threads]
#include "../amplisolve_b200/csrc/as_host.cpp"

int main(int argc, char** argv) {
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
