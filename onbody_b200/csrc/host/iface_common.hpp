// iface_common.hpp - shared plumbing of the two drop-in shim libraries (host C++ over the C ABI only).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "onbody_b200.h"

namespace onb_shim {

inline float fail(onb_context* c, const char* where) {
    std::fprintf(stderr, "onbody_b200 (%s): %s\n", where, c ? onb_error(c) : onb_last_create_error());
    const char* mode = std::getenv("ONBODY_B200_ON_ERROR");
    if (mode && std::strcmp(mode, "return") == 0) return -1.0f;
    std::abort();
}

// one lazily-created context per physics, reused across calls (time-stepping callers call these every step)
inline onb_context* context(int physics, float /*unused*/ = 0.f) {
    static onb_context* ctx[8] = {nullptr};
    if (!ctx[physics]) {
        int dev = 0;
        if (const char* d = std::getenv("ONBODY_B200_DEVICE")) dev = std::atoi(d);
        ctx[physics] = onb_create(physics, dev);
        if (ctx[physics] && onb_set_params(ctx[physics], 128, 4, ONB_ARITH_FAST) != ONB_OK) return nullptr;   // blockSize 128, order 4
    }
    return ctx[physics];
}

// solver = makeTree(srcs) -> barycentric upward -> makeTree(targs) -> boxwise treecode -> += in original order
// direct = nbody_naive -> += (targets were never reordered)
// The caller's separate coordinate / strength / output arrays go to the library as they are (one pointer per plane): no
// re-packing on the host, and the += of the solver path is one scatter kernel + pipelined pinned read-back (or fully in place
// when the caller's arrays are device memory).
inline float run(int physics, bool direct, float theta, int ns, const float* const* sx, int PD, const float* const* ss, int SD, const float* sr,
                 int nt, const float* const* tx, const float* tr, float* const* out, int OD) {
    (void)PD; (void)SD;
    onb_context* c = context(physics);
    if (!c) return fail(nullptr, "create");
    if (physics == ONB_VORT2DTR) onb_set_flops_per_pair(c, 13);      // interface2dvorttr.cpp:53
    std::vector<float> r0;
    const float* trp = tr;
    if (!trp) { r0.assign(nt, 0.0f); trp = r0.data(); }        // the reference leaves targs.r zero-initialised here
    if (onb_set_sources_planes(c, (uint64_t)ns, sx, sr, ss) != ONB_OK) return fail(c, "set_sources");
    if (onb_set_targets_planes(c, (uint64_t)nt, tx, trp) != ONB_OK) return fail(c, "set_targets");
    float flops = 0.0f;
    if (direct) {
        std::vector<float> u((size_t)OD * nt, 0.0f);
        if (onb_zero_vels(c) != ONB_OK || onb_naive(c, 1, &flops) != ONB_OK) return fail(c, "naive");
        if (onb_get_parts(c, 1, nullptr, nullptr, nullptr, u.data(), nullptr) != ONB_OK) return fail(c, "get");
        for (int d = 0; d < OD; ++d) { float* o = out[d]; const float* ud = u.data() + (size_t)d * nt; for (int i = 0; i < nt; ++i) o[i] += ud[i]; }
    } else {
        if (onb_make_trees(c) != ONB_OK) return fail(c, "make_trees");         // both builds overlap on the device
        if (onb_upward(c, 0) != ONB_OK) return fail(c, "upward");
        if (onb_zero_vels(c) != ONB_OK || onb_treecode3(c, theta, &flops) != ONB_OK) return fail(c, "treecode3");
        if (onb_add_results_planes(c, out) != ONB_OK) return fail(c, "results");   // interface3dvortgrads.cpp:384-395
    }
    return flops;
}

}  // namespace onb_shim
