// bench_similarities.cpp — C++ twin of the GPU blocks of StringWars `similarities/bench.rs`, driving the
// stringwars_b200 contender through its C ABI exactly the way the Rust harness drives `stringzillas` engines.
//
// The reference is compiled Rust; this image has no Rust toolchain, so the compiled host side is C++. Mirrors:
//   env knobs            STRINGWARS_DATASET/_TOKENS/_MAX_TOKENS (utils.rs:273-433), _TIME/_WARMUP (utils.rs:574-576),
//                        _FILTER (utils.rs:457-483), _BATCH_PER_CORE (utils.rs:806-808)
//   batch sizing         auto_batch_size(SMs, 256) and crossproduct_side (bench.rs:93, 113-117, 286-289)
//   accounting           cells = sum|q| * sum|c|, bytes = sum|q| + sum|c| (bench.rs:216-224), CUPS = cells / wall time
//   loop                 measure_throughput: uncounted warm-up, wall-time budget, >= 1 measured call (utils.rs:721-799)
//   variant names        uniform/…<{n}gpu>, linear/…, affine/… (bench.rs:330-363, 591)
//   pre-flight           verify-then-bench in the style of containers/bench.rs:344-357
// No CPU contenders here: the product has no CPU path; the CPU baselines live in bench.py (oracle restatement).
//
// Build: make -C harness      Run: STRINGWARS_TIME=1 harness/bench_similarities   (needs a CUDA device)
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <regex>
#include <sstream>
#include <string>
#include <vector>

#include "../include/stringwars_b200.h"

using Clock = std::chrono::steady_clock;

static double env_f64(const char *name, double fallback) {
    const char *v = std::getenv(name);
    return v ? std::atof(v) : fallback;
}
static size_t env_usize(const char *name, size_t fallback) {
    const char *v = std::getenv(name);
    return v && std::atoll(v) > 0 ? (size_t)std::atoll(v) : fallback;
}

struct BenchBudget { double warmup_s, time_s; };
struct WorkUnits { uint64_t elements, bytes; };

static bool should_run(const std::string &name) {
    const char *pattern = std::getenv("STRINGWARS_FILTER");
    if (!pattern || !*pattern) return true;
    try { return std::regex_search(name, std::regex(pattern)); } catch (...) { return true; }
}

static std::string scale_si(double v) {
    char buf[64];
    if (v >= 1e12) std::snprintf(buf, sizeof buf, "%.2f T", v / 1e12);
    else if (v >= 1e9) std::snprintf(buf, sizeof buf, "%.2f G", v / 1e9);
    else if (v >= 1e6) std::snprintf(buf, sizeof buf, "%.2f M", v / 1e6);
    else if (v >= 1e3) std::snprintf(buf, sizeof buf, "%.2f k", v / 1e3);
    else std::snprintf(buf, sizeof buf, "%.2f ", v);
    return buf;
}

template <typename F>
static void measure_throughput(const std::string &name, const BenchBudget &budget, F &&closure) {
    if (!should_run(name)) return;
    auto t0 = Clock::now();
    while (budget.warmup_s > 0 && std::chrono::duration<double>(Clock::now() - t0).count() < budget.warmup_s) closure();
    uint64_t elements = 0, bytes = 0, calls = 0;
    std::vector<double> lat_us;
    t0 = Clock::now();
    for (;;) {
        auto c0 = Clock::now();
        WorkUnits u = closure();
        auto c1 = Clock::now();
        lat_us.push_back(std::chrono::duration<double, std::micro>(c1 - c0).count());
        elements += u.elements, bytes += u.bytes, ++calls;
        if (std::chrono::duration<double>(c1 - t0).count() >= budget.time_s) break;
    }
    const double secs = std::chrono::duration<double>(Clock::now() - t0).count();
    std::sort(lat_us.begin(), lat_us.end());
    auto q = [&](double f) { return lat_us[std::min(lat_us.size() - 1, (size_t)(f * lat_us.size()))]; };
    std::printf("%-50s %sCUPS | %sB/s | p50 %.1f us | p99 %.1f us | %llu calls\n", name.c_str(), scale_si(elements / secs).c_str(),
                scale_si(bytes / secs).c_str(), q(0.5), q(0.99), (unsigned long long)calls);
    std::fflush(stdout);
}

// Arrow-style tape in CUDA managed memory — the reference's own `BytesTape<u64, UnifiedAlloc>` (bench.rs:292-295): the
// library reads it in place, no staging copy. STRINGWARS_B200_MEMORY=pinned switches tape and matrix to page-locked host
// memory (H2D / D2H inside every call).
struct Tape {
    uint8_t *data = nullptr;
    uint64_t *offsets = nullptr;
    size_t count = 0;
    size_t len(size_t i) const { return offsets[i + 1] - offsets[i]; }
};

static std::vector<std::string> load_tokens() {
    std::vector<std::string> tokens;
    const char *path = std::getenv("STRINGWARS_DATASET");
    const std::string mode = std::getenv("STRINGWARS_TOKENS") ? std::getenv("STRINGWARS_TOKENS") : "words";
    const size_t max_tokens = env_usize("STRINGWARS_MAX_TOKENS", SIZE_MAX);
    if (path && *path) {
        std::ifstream in(path, std::ios::binary);
        if (!in) { std::fprintf(stderr, "cannot open STRINGWARS_DATASET=%s\n", path); std::exit(2); }
        std::stringstream ss; ss << in.rdbuf();
        const std::string text = ss.str();
        size_t start = 0;
        for (size_t i = 0; i <= text.size() && tokens.size() < max_tokens; ++i) {
            const bool sep = i == text.size() || text[i] == '\n' || (mode == "words" && text[i] == ' ');
            if (!sep) continue;
            if (i > start) tokens.emplace_back(text.substr(start, i - start));
            start = i + 1;
        }
    } else { // no dataset on the box: seeded synthetic words (a-z, 1..16 bytes), seed 42 like the reference's default
        uint64_t state = 42;
        auto next = [&] { state ^= state << 13; state ^= state >> 7; state ^= state << 17; return state; };
        const size_t n = std::min<size_t>(max_tokens, 200000);
        for (size_t i = 0; i < n; ++i) {
            std::string w(1 + next() % 16, 'a');
            for (auto &ch : w) ch = (char)('a' + next() % 26);
            tokens.push_back(std::move(w));
        }
    }
    return tokens;
}

#define CHECK(call)                                                          \
    do {                                                                     \
        const char *err_ = nullptr;                                          \
        if ((call) != SWB_OK) { std::fprintf(stderr, "%s\n  -> %s\n", #call, err_ ? err_ : "?"); std::exit(1); } \
    } while (0)

int main() {
    const BenchBudget budget{env_f64("STRINGWARS_WARMUP", 5.0), env_f64("STRINGWARS_TIME", 30.0)}; // bench.rs:1031
    const char *err = nullptr;
    if (swb_device_count() == 0) { std::printf("no CUDA device: every stringwars_b200 variant SKIPPED\n"); return 0; }

    std::vector<std::string> tokens = load_tokens();
    if (tokens.size() < 2) { std::fprintf(stderr, "Dataset must contain at least two items for comparisons.\n"); return 2; }
    Tape tape;
    size_t total = 0;
    for (auto &t : tokens) total += t.size();
    const bool pinned = std::getenv("STRINGWARS_B200_MEMORY") && std::string(std::getenv("STRINGWARS_B200_MEMORY")) == "pinned";
    auto alloc = pinned ? swb_host_alloc : swb_unified_alloc;
    auto release = pinned ? swb_host_free : swb_unified_free;
    if (alloc(total + 64, (void **)&tape.data, &err) != SWB_OK || alloc((tokens.size() + 1) * 8, (void **)&tape.offsets, &err) != SWB_OK) {
        std::fprintf(stderr, "tape allocation failed: %s\n", err); return 1;
    }
    tape.count = tokens.size();
    tape.offsets[0] = 0;
    for (size_t i = 0; i < tokens.size(); ++i) {
        std::memcpy(tape.data + tape.offsets[i], tokens[i].data(), tokens[i].size());
        tape.offsets[i + 1] = tape.offsets[i] + tokens[i].size();
    }

    int sms = 64;
    swb_multiprocessor_count(0, &sms, &err);
    const size_t batch_gpu = (size_t)sms * env_usize("STRINGWARS_BATCH_PER_CORE", 256); // auto_batch_size (utils.rs:815-819)
    const size_t side = std::max<size_t>(1, std::min<size_t>((size_t)std::llround(std::sqrt((double)batch_gpu)), tape.count / 2));
    std::printf("Benchmark configuration:\n- GPU batch: %zu (%zux%zu cross-product)\n- Tokens available: %zu\n\n", batch_gpu, side, side, tape.count);

    uint64_t sum_q = 0, sum_c = 0; // crossproduct_metrics_bytes (bench.rs:216-224)
    for (size_t i = 0; i < side; ++i) sum_q += tape.len(i), sum_c += tape.len(side + i);
    const uint64_t cells = sum_q * sum_c, bytes = sum_q + sum_c;
    // crossproduct_metrics_chars (bench.rs:230-247): the UTF-8 engine's cells are counted in code points, i.e. bytes that
    // are not continuation bytes (10xxxxxx)
    auto chars_of = [&](size_t i) {
        uint64_t n = 0;
        for (uint64_t p = tape.offsets[i]; p < tape.offsets[i + 1]; ++p) n += (tape.data[p] & 0xC0) != 0x80;
        return n;
    };
    uint64_t chars_q = 0, chars_c = 0;
    for (size_t i = 0; i < side; ++i) chars_q += chars_of(i), chars_c += chars_of(side + i);
    const uint64_t cells_chars = chars_q * chars_c;
    const uint64_t *q_offs = tape.offsets, *c_offs = tape.offsets + side; // zero-copy subviews [0, side) and [side, 2*side)

    uint64_t *matrix = nullptr;
    if (alloc(side * side * sizeof(uint64_t), (void **)&matrix, &err) != SWB_OK) { std::fprintf(stderr, "%s\n", err); return 1; } // UnifiedMat (bench.rs:130)

    uint8_t byte_to_class[256];
    for (int i = 0; i < 256; ++i) byte_to_class[i] = (uint8_t)(i % 32); // unary_class_costs (bench.rs:98-108)
    int8_t costs[32 * 32];
    for (int i = 0; i < 32; ++i) for (int j = 0; j < 32; ++j) costs[i * 32 + j] = i == j ? 2 : -1;

    const int devices = swb_device_count();
    for (int gpus : {1, 2, 4, 8}) {
        if (gpus > devices) break;
        swb_scope_t *scope = nullptr;
        if (swb_scope_create(0, gpus, &scope, &err) != SWB_OK) { std::fprintf(stderr, "<%dgpu>: SKIPPED (%s)\n", gpus, err); continue; }
        const std::string tag = "<" + std::to_string(gpus) + "gpu>";
        if (gpus == 1) std::printf("# uniform\n");
        {
            swb_engine_t *lev = nullptr;
            if (swb_levenshtein_create(scope, 0, 1, 1, 1, &lev, &err) == SWB_OK) {
                // pre-flight: the cross product must agree with the scalar surface on its diagonal before it is timed
                CHECK(swb_compute_cross(lev, tape.data, q_offs, side, tape.data, c_offs, side, matrix, side, &err_));
                for (size_t i = 0; i < std::min<size_t>(side, 32); ++i) {
                    const uint64_t d = swb_levenshtein_pair(lev, tape.data + q_offs[i], tape.len(i), tape.data + c_offs[i], tape.len(side + i), &err);
                    if (d != matrix[i * side + i]) { std::fprintf(stderr, "pre-flight mismatch on pair %zu: %llu vs %llu\n", i, (unsigned long long)d, (unsigned long long)matrix[i * side + i]); return 3; }
                }
                measure_throughput("uniform/stringwars_b200::levenshtein" + tag, budget, [&] {
                    CHECK(swb_compute_cross(lev, tape.data, q_offs, side, tape.data, c_offs, side, matrix, side, &err_));
                    return WorkUnits{cells, bytes};
                });
                swb_engine_destroy(lev);
            }
            swb_engine_t *lev8 = nullptr;
            if (swb_levenshtein_utf8_create(scope, 0, 1, 1, 1, &lev8, &err) == SWB_OK) {
                // pre-flight: code-point distances are bounded by the two lengths in code points and, on the diagonal of a
                // pure-ASCII batch, equal the byte engine's
                CHECK(swb_compute_cross(lev8, tape.data, q_offs, side, tape.data, c_offs, side, matrix, side, &err_));
                for (size_t i = 0; i < std::min<size_t>(side, 32); ++i) {
                    const uint64_t d = matrix[i * side + i], ca = chars_of(i), cb = chars_of(side + i);
                    if (d > std::max(ca, cb) || d < (ca > cb ? ca - cb : cb - ca)) { std::fprintf(stderr, "utf8 pre-flight: pair %zu distance %llu outside [%llu, %llu]\n", i, (unsigned long long)d, (unsigned long long)(ca > cb ? ca - cb : cb - ca), (unsigned long long)std::max(ca, cb)); return 3; }
                    if (chars_q == sum_q && chars_c == sum_c) {
                        const uint64_t b = swb_levenshtein_pair(lev8, tape.data + q_offs[i], tape.len(i), tape.data + c_offs[i], tape.len(side + i), &err);
                        if (b != d) { std::fprintf(stderr, "utf8 pre-flight mismatch on pair %zu\n", i); return 3; }
                    }
                }
                measure_throughput("uniform/stringwars_b200::levenshtein_utf8" + tag, budget, [&] {
                    CHECK(swb_compute_cross(lev8, tape.data, q_offs, side, tape.data, c_offs, side, matrix, side, &err_));
                    return WorkUnits{cells_chars, bytes}; // crossproduct_metrics_chars (bench.rs:612)
                });
                swb_engine_destroy(lev8);
            }
        }
        if (gpus == 1) {
            struct Group { const char *name; int open, extend; };
            for (Group g : {Group{"linear", -2, -2}, Group{"affine", -5, -1}}) { // bench.rs:342-362
                std::printf("# %s\n", g.name);
                swb_engine_t *nw = nullptr, *sw = nullptr;
                if (swb_needleman_wunsch_create(scope, byte_to_class, costs, g.open, g.extend, &nw, &err) == SWB_OK) {
                    measure_throughput(std::string(g.name) + "/stringwars_b200::needleman_wunsch" + tag, budget, [&] {
                        CHECK(swb_compute_cross(nw, tape.data, q_offs, side, tape.data, c_offs, side, matrix, side, &err_));
                        return WorkUnits{cells, bytes};
                    });
                    swb_engine_destroy(nw);
                }
                if (swb_smith_waterman_create(scope, byte_to_class, costs, g.open, g.extend, &sw, &err) == SWB_OK) {
                    measure_throughput(std::string(g.name) + "/stringwars_b200::smith_waterman" + tag, budget, [&] {
                        CHECK(swb_compute_cross(sw, tape.data, q_offs, side, tape.data, c_offs, side, matrix, side, &err_));
                        return WorkUnits{cells, bytes};
                    });
                    swb_engine_destroy(sw);
                }
            }
        }
        swb_scope_destroy(scope);
    }
    uint64_t checksum = 0;
    for (size_t i = 0; i < side * side; ++i) checksum += matrix[i];
    std::fprintf(stderr, "  last matrix checksum=%llu\n", (unsigned long long)checksum); // cf. bench.py:305, 622
    release(matrix), release(tape.data), release(tape.offsets);
    return 0;
}
