#include "pnn_hm_direct.h"

#include <errno.h>    // program_invocation_short_name (glibc)

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>

namespace {

pnn_handle* g_handle = NULL;
long long g_calls[5] = {0, 0, 0, 0, 0};
double g_seconds[5] = {0., 0., 0., 0., 0.};
double g_first_seconds[5] = {0., 0., 0., 0., 0.};   // the first call of a width: device initialisation (first width) and the upload of the net
long long g_posts[5] = {0, 0, 0, 0, 0};             // requests posted ahead of their use (prefetch)
bool g_first_done[5] = {false, false, false, false, false};
bool g_cache_enabled = true;   // without the memo an answer that is not collected by the very next call would be lost: nothing is posted ahead

int width_index(int width) { return width == 4 ? 0 : width == 8 ? 1 : width == 16 ? 2 : width == 32 ? 3 : 4; }

// host time spent inside the library for one width: posting and collecting both count
void account(int index, double dt) {
    if (!g_first_done[index]) {
        g_first_done[index] = true;
        g_first_seconds[index] = dt;
    } else {
        g_seconds[index] += dt;
    }
}

void print_stats() {
    const char* path = getenv("PNN_HM_STATS");
    FILE* f = path ? fopen(path, "w") : stderr;
    if (!f) f = stderr;
    long long total(0);
    double seconds(0.);
    for (int i(0); i < 5; i++) {
        // the first call of a width waits for the device initialisation / the upload of its net: counted in the total, kept out
        // of the per-call figure
        const long long later(g_calls[i] > 1 ? g_calls[i] - 1 : 0);
        fprintf(f, "pnn_calls width %d: %lld calls, %.6f s, %.2f us/call (first call %.1f ms, not in the per-call figure; %lld posted ahead)\n", 4 << i,
                g_calls[i], g_seconds[i] + g_first_seconds[i], later ? 1.e6 * g_seconds[i] / later : 0., 1.e3 * g_first_seconds[i], g_posts[i]);
        total += g_calls[i];
        seconds += g_seconds[i] + g_first_seconds[i];
    }
    fprintf(f, "pnn_calls total: %lld calls, %.6f s\n", total, seconds);
    if (g_handle) {
        int64_t hits(0), misses(0);
        pnn_hm_cache_stats(g_handle, &hits, &misses);
        fprintf(f, "pnn_cache: %lld hits, %lld misses\n", (long long)hits, (long long)misses);
    }
    if (f != stderr) fclose(f);
    if (g_handle) {
        pnn_release_at_exit(g_handle);      // the process ends here: no buffer-by-buffer teardown
        g_handle = NULL;
    }
}

}  // namespace

namespace pnn_hm_direct {

float read_mean_file(const std::string& path_to_mean_training) {
    std::ifstream file(path_to_mean_training.c_str(), std::ios::binary);
    if (!file) {
        fprintf(stderr, "The file at \"%s\" cannot be opened.\n", path_to_mean_training.c_str());
        abort();
    }
    const std::string data((std::istreambuf_iterator<char>(file)), std::istreambuf_iterator<char>());
    std::size_t pos(0);
    if (data.size() >= 2 && static_cast<unsigned char>(data[0]) == 0x80) pos = 2;       // PROTO opcode
    if (pos < data.size() && data[pos] == 'G' && pos + 9 <= data.size()) {             // BINFLOAT: 8 big-endian bytes
        uint64_t bits(0);
        for (int i(0); i < 8; i++) bits = (bits << 8) | static_cast<unsigned char>(data[pos + 1 + i]);
        double value;
        memcpy(&value, &bits, 8);
        return static_cast<float>(value);
    }
    if (pos < data.size() && data[pos] == 'F') pos += 1;                                 // FLOAT (text)
    char* end(NULL);
    const double value(strtod(data.c_str() + pos, &end));
    if (end == data.c_str() + pos) {
        fprintf(stderr, "The file at \"%s\" does not hold a float.\n", path_to_mean_training.c_str());
        abort();
    }
    return static_cast<float>(value);
}

pnn_handle* create(const std::string& path_to_file_paths_to_graphs_output, float mean_training, int qp_selection) {
    if (g_handle) return g_handle;                      // one TComPrediction per HM process
    int device(0);
    const char* env = getenv("PNN_DEVICE");
    if (env) device = atoi(env);
    if (pnn_create_deferred(path_to_file_paths_to_graphs_output.c_str(), mean_training, qp_selection, device, &g_handle) != 0) {
        fprintf(stderr, "%s\n", pnn_last_error(NULL));
        return NULL;
    }
    const char* cache = getenv("PNN_HM_CACHE");
    g_cache_enabled = cache ? atoi(cache) != 0 : true;
    pnn_set_hm_cache(g_handle, g_cache_enabled ? 1 : 0);
    // HM does not touch the reconstruction between initIntraPatternChType and predIntraAng: the context is copied only
    // when the neural-network mode is actually evaluated
    pnn_set_context_lazy(g_handle, 1);
    // An encoder will need every net, but not before it has coded its first row of coding tree units (no causal context
    // there): device initialisation and uploads start now, on a thread of the library.  A decoder may never need them.
    const char* warm = getenv("PNN_HM_WARM_UP");
    const bool is_encoder(program_invocation_short_name && strstr(program_invocation_short_name, "Encoder") != NULL);
    if (warm ? atoi(warm) != 0 : is_encoder) pnn_warm_up(g_handle);
    atexit(print_stats);
    return g_handle;
}

int set_context(pnn_handle* handle, int width, const int* piRoiOrigin, int iPicStride, const bool* bNeighborFlags,
                int iNumIntraNeighbor, int iUnitWidth, int iUnitHeight, int iAboveUnits, int iLeftUnits) {
    uint8_t flags[2 * 64 + 1];
    const int total(iAboveUnits + iLeftUnits + 1);
    if (bNeighborFlags && total > 0 && total <= 129) {
        for (int i(0); i < total; i++) flags[i] = bNeighborFlags[i] ? 1 : 0;
    }
    const int code(pnn_set_context(handle, width, piRoiOrigin, iPicStride, bNeighborFlags ? flags : NULL, iNumIntraNeighbor,
                                   iUnitWidth, iUnitHeight, iAboveUnits, iLeftUnits));
    if (code != 0) fprintf(stderr, "%s\n", pnn_last_error(handle));
    return code;
}

int prefetch(pnn_handle* handle, int width) {
    static const bool enabled(getenv("PNN_HM_PREFETCH") ? atoi(getenv("PNN_HM_PREFETCH")) != 0 : true);
    if (!enabled || !g_cache_enabled) return 0;
    const std::chrono::steady_clock::time_point t0(std::chrono::steady_clock::now());
    const int code(pnn_predict_hm_begin(handle, width));
    const double dt(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    if (code != 0) {
        fprintf(stderr, "%s\n", pnn_last_error(handle));
        return code;
    }
    const int index(width_index(width));
    account(index, dt);
    g_posts[index] += 1;
    return 0;
}

int prefetch_first_quadrant(pnn_handle* handle, int width) {
    static const bool enabled(getenv("PNN_HM_PREFETCH_QUADRANT") ? atoi(getenv("PNN_HM_PREFETCH_QUADRANT")) != 0 : true);
    if (!enabled) return 0;
    return prefetch(handle, width);
}

int predict(pnn_handle* handle, int width, int* piPred, int stride) {
    const std::chrono::steady_clock::time_point t0(std::chrono::steady_clock::now());
    const int code(pnn_predict_hm(handle, width, piPred, stride));
    const double dt(std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    if (code != 0) {
        fprintf(stderr, "%s\n", pnn_last_error(handle));
        return code;
    }
    const int index(width_index(width));
    account(index, dt);
    g_calls[index] += 1;
    return 0;
}

}  // namespace pnn_hm_direct
