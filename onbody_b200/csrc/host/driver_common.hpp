/*
 * driver_common.hpp - the reference's command-line drivers (src/ongrav3d.cpp:465-912, onvort3d.cpp, onvort2d.cpp,
 * onvortgrad3d.cpp main()) with the L0-L2 template calls replaced by the C ABI of the CUDA library.
 * Same flags (-n= -t= -t1..4= -o= -b= -h), same defaults, same stdout grammar (scripts/speedtest.pl greps
 * "error in fastsumm", "fast total", "onbody naive"), same seeded inputs, same error metric. Host C++ only.
 *
 * Deliberately kept quirks: "-t1=".."-t4=" parse atof(argv+3), i.e. from the '=' sign, get 0 and print the usage
 * (ongrav3d.cpp:491-506); -n goes through atoi. With -o omitted (order = -1, the reference's default) the legacy pair-merge
 * equivalents are used for treecode2/3 exactly like the reference (refineTree(srcs) + calcEquivalents); the dual tree is
 * skipped in that mode with a note on stderr, because the reference builds no target equivalents there at all
 * (calcEquivalents returns at barneshut.hpp:953 for targets). Extra flags: -strict selects ARITH_STRICT; -g=<n> (or
 * ONBODY_B200_GPUS=<n>) runs on n GPUs of this node: one context and one host thread per GPU, the library's NCCL
 * communicator (onb_comm_init_all) behind the very same phase calls, every rank filling its target shard of the result
 * arrays; -lean selects the lean memory mode (N = 1e9 on 8 GPUs). Output is identical for every GPU count.
 * The direct-sum sample uses the divisor of the reference's OpenMP non-Vc build (ongrav3d.cpp:560), the build the oracle uses.
 */
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>
#include "onbody_b200.h"

struct DriverSpec {
    const char* progname; int physics; int PD, SD, OD; int strength_mode;
    bool multi_theta;      // ongrav3d/onvort3d/onvort2d have theta1..4; onvortgrad3d a single theta
    bool has_fast;         // test_iterations[4]
    const char* banner;    // extra init line (ongrav3d prints the charges line), or nullptr
    float t1, t2, t3, t4;  // default thetas
};

static const DriverSpec* g_spec = nullptr;
static void usage() {
    std::fprintf(stderr, "Usage: %s [-h] [-n=<nparticles>] [-t=<theta>] [-o=<order>] [-b=<blocksize>]\n", g_spec->progname);
    std::exit(1);
}
#define DRV_CHECK(call) do { int rc__ = (call); if (rc__ != ONB_OK) { std::fprintf(stderr, "%s: %s failed: %s\n", g_spec->progname, #call, onb_error(ctx)); std::exit(2); } } while (0)
// one phase on every GPU: f(rank) on one host thread per context (the collectives inside the library rendezvous across them)
template <class F> static void on_all(int ngpus, F f) {
    if (ngpus == 1) { f(0); return; }
    std::vector<std::thread> th;
    for (int i = 0; i < ngpus; ++i) th.emplace_back(f, i);
    for (auto& t : th) t.join();
}

static inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int run_driver(int argc, char* argv[], const DriverSpec& spec) {
    g_spec = &spec;
    const size_t minBlkSz = 2;
    size_t numSrcs = 10000, numTargs = 10000;
    size_t blockSize = minBlkSz * ((128 + minBlkSz - 1) / minBlkSz);
    size_t eqBlockSize = blockSize;
    const size_t echonum = 1;
    float theta1 = spec.t1, theta2 = spec.t2, theta3 = spec.t3, theta4 = spec.t4;
    int order = -1, arith = ONB_ARITH_FAST;
    int ngpus = std::getenv("ONBODY_B200_GPUS") ? std::max(1, atoi(std::getenv("ONBODY_B200_GPUS"))) : 1;
    bool lean = false;
    for (int i = 1; i < argc; i++) {
        if (strncmp(argv[i], "-n=", 3) == 0) {
            size_t num = atoi(argv[i] + 3); if (num < 1) usage(); numSrcs = num; numTargs = num;
        } else if (spec.multi_theta && strncmp(argv[i], "-t1=", 4) == 0) { float t = atof(argv[i] + 3); if (t < 0.0001) usage(); theta1 = t;
        } else if (spec.multi_theta && strncmp(argv[i], "-t2=", 4) == 0) { float t = atof(argv[i] + 3); if (t < 0.0001) usage(); theta2 = t;
        } else if (spec.multi_theta && strncmp(argv[i], "-t3=", 4) == 0) { float t = atof(argv[i] + 3); if (t < 0.0001) usage(); theta3 = t;
        } else if (spec.multi_theta && strncmp(argv[i], "-t4=", 4) == 0) { float t = atof(argv[i] + 3); if (t < 0.0001) usage(); theta4 = t;
        } else if (strncmp(argv[i], "-t=", 3) == 0) {
            float t = atof(argv[i] + 3); if (t < 0.0001) usage(); theta1 = theta2 = theta3 = theta4 = t;
        } else if (strncmp(argv[i], "-o=", 3) == 0) {
            int o = atoi(argv[i] + 3); if (o < 1) usage(); order = o;
        } else if (strncmp(argv[i], "-b=", 3) == 0) {
            size_t num = atoi(argv[i] + 3); if (num < 1) usage();
            blockSize = minBlkSz * ((num + minBlkSz - 1) / minBlkSz); eqBlockSize = blockSize;
        } else if (strcmp(argv[i], "-strict") == 0) { arith = ONB_ARITH_STRICT;
        } else if (strncmp(argv[i], "-g=", 3) == 0) { ngpus = std::max(1, atoi(argv[i] + 3));
        } else if (strcmp(argv[i], "-lean") == 0) { lean = true;
        } else if (strncmp(argv[i], "-h", 2) == 0 || strncmp(argv[i], "--h", 3) == 0) usage();
    }
    std::string withwhat;
    const bool legacy = order < 0;
    if (legacy) {
        withwhat = "equivalent particles";                                               // ongrav3d.cpp:543-544
    } else {
        withwhat = "a barycentric grid";
        eqBlockSize = 128;     // the GPU build pads (order+1)^PD to one 128-slot block (the reference pads to its SIMD width)
    }

    std::printf("Running %s with %ld sources and %ld targets\n", spec.progname, (long)numSrcs, (long)numTargs);
    std::printf("  source block sizes %ld and %ld, target block size %ld\n\n", (long)blockSize, (long)eqBlockSize, (long)blockSize);
    size_t ntskip = std::max(1, (int)((float)numSrcs * (float)numTargs / 2.e+9));

    const int dev0 = std::getenv("ONBODY_B200_DEVICE") ? atoi(std::getenv("ONBODY_B200_DEVICE")) : 0;
    // verification aid: ONBODY_B200_LOOPBACK=1 puts all -g contexts on ONE device and joins them with the library's loopback
    // transport (device copies instead of NCCL), so the multi-GPU driver path can be checked where only one GPU exists
    const bool loopback = std::getenv("ONBODY_B200_LOOPBACK") != nullptr;
    std::vector<onb_context*> ctxs(ngpus, nullptr);
    for (int g = 0; g < ngpus; ++g) {
        ctxs[g] = onb_create(spec.physics, loopback ? dev0 : dev0 + g);
        if (!ctxs[g]) { std::fprintf(stderr, "%s: GPU %d: %s\n", spec.progname, dev0 + g, onb_last_create_error()); return 2; }
        onb_context* ctx = ctxs[g];
        DRV_CHECK(onb_set_params(ctx, (int)blockSize, order, arith));
        if (lean) DRV_CHECK(onb_set_memory_mode(ctx, ONB_MEM_LEAN));
    }
    onb_context* ctx = ctxs[0];
    if (ngpus > 1) {
        if (legacy) { std::fprintf(stderr, "%s: -g needs barycentric equivalents (-o=<order>)\n", spec.progname); return 2; }
        if (loopback) DRV_CHECK(onb_comm_init_loopback(ctxs.data(), ngpus)); else DRV_CHECK(onb_comm_init_all(ctxs.data(), ngpus));
        std::fprintf(stderr, "%s: %d %s, targets sharded by leaf range, %s communicator inside the library\n", spec.progname, ngpus,
                     loopback ? "contexts on one GPU" : "GPUs", loopback ? "loopback" : "NCCL");
    }
    // ctx is rank 0 below; ALL(call) makes the same call on every rank
#define ALL(call) on_all(ngpus, [&](int rk__) { onb_context* ctx = ctxs[rk__]; DRV_CHECK(call); })

    std::printf("Allocate and initialize\n");
    double start = now_s();
    std::vector<float> x((size_t)spec.PD * numSrcs), r(numSrcs), s((size_t)spec.SD * numSrcs);
    onb_driver_inputs(spec.physics, numSrcs, spec.strength_mode, x.data(), r.data(), s.data());
    if (spec.banner) std::printf("%s\n", spec.banner);
    if (ngpus > 1) ALL(onb_set_sliced_inputs(ctx, 1));                 // every GPU pulls 1/ngpus of the planes over PCIe, NVLink replicates
    ALL(onb_set_sources(ctx, numSrcs, x.data(), r.data(), s.data()));
    ALL(onb_set_targets(ctx, numTargs, x.data(), r.data()));           // the drivers copy the engine: same positions (ongrav3d.cpp:574-594)
    std::printf("  init parts time:\t\t[%.4f] seconds\n", now_s() - start);
    std::vector<double> treetime(5, 0.0);

    std::printf("\nBuilding the source tree\n");
    std::printf("  with %ld particles and block size of %ld\n", (long)numSrcs, (long)blockSize);
    start = now_s(); ALL(onb_make_tree(ctx, 0)); double dt = now_s() - start;
    std::printf("  build tree time:\t\t[%.4f] seconds\n", dt);
    for (int k = 1; k < 5; ++k) treetime[k] += dt;
    if (legacy) {                                                                        // ongrav3d.cpp:617-630
        start = now_s(); ALL(onb_refine(ctx, 0)); dt = now_s() - start;
        std::printf("  refine within leaf nodes:\t[%.4f] seconds\n", dt);
        for (int k = 2; k < 5; ++k) treetime[k] += dt;
    }
    std::printf("  add buffer at end of srcs:\t[%.4f] seconds\n", 0.0);
    std::printf("\nCalculating equivalent particles\n");
    int levels = 0, numnodes = 0; onb_tree_shape(ctx, 0, &levels, &numnodes);
    std::printf("  need %ld particles and block size of %ld\n", (long)(numnodes / 2) * (long)eqBlockSize, (long)eqBlockSize);
    std::printf("  allocate eqsrcs structures:\t[%.4f] seconds\n", 0.0);
    start = now_s(); ALL(onb_upward(ctx, 0)); dt = now_s() - start;
    std::printf(legacy ? "  create equivalent parts:\t[%.4f] seconds\n" : "  create barylagrange parts:\t[%.4f] seconds\n", dt);
    for (int k = 2; k < 5; ++k) treetime[k] += dt;

    std::printf("\nBuilding the target tree\n");
    std::printf("  with %ld particles and block size of %ld\n", (long)numTargs, (long)blockSize);
    start = now_s(); ALL(onb_make_tree(ctx, 1)); dt = now_s() - start;
    std::printf("  build tree time:\t\t[%.4f] seconds\n", dt);
    treetime[3] += dt; treetime[4] += dt;
    const bool run_fast = spec.has_fast && !legacy;
    if (spec.has_fast && legacy)
        std::fprintf(stderr, "%s: -o omitted: the dual-tree method is skipped (the reference's legacy path builds no target equivalents, "
                             "barneshut.hpp:953); pass -o=<order> for it\n", spec.progname);
    if (spec.has_fast) {
        // (also with -o omitted: the reference refines the target leaves whenever the dual tree is scheduled,
        // ongrav3d.cpp:686-724, which fixes the target order every method's "particle 0" line refers to)
        std::printf("\nCalculating equivalent targ points\n");
        onb_tree_shape(ctx, 1, &levels, &numnodes);
        std::printf("  need %ld particles and block size of %ld\n", (long)(numnodes / 2) * (long)eqBlockSize, (long)eqBlockSize);
        std::printf("  allocate eqtargs structures:\t[%.4f] seconds\n", 0.0);
        start = now_s(); ALL(onb_refine(ctx, 1)); dt = now_s() - start;
        std::printf("  refine within leaf nodes:\t[%.4f] seconds\n", dt); treetime[4] += dt;
        start = now_s(); if (run_fast) ALL(onb_upward(ctx, 1)); dt = now_s() - start;
        std::printf("  create equivalent parts:\t[%.4f] seconds\n", dt); treetime[4] += dt;
    }

    std::vector<float> u((size_t)spec.OD * numTargs), naiveu(numTargs);
    float flops = 0.0f;
    std::vector<float> rank_flops(ngpus, 0.0f);
    auto sum_flops = [&]() { flops = 0.0f; for (float f : rank_flops) flops += f; };
    auto fetch = [&]() {
        if (ngpus == 1) { DRV_CHECK(onb_get_parts(ctx, 1, nullptr, nullptr, nullptr, u.data(), nullptr)); return; }
        ALL(onb_get_shard_results(ctx, u.data(), 0, nullptr, nullptr));      // every rank fills its own target range
    };
    auto echo = [&]() {
        for (size_t i = 0; i < echonum * ntskip; i += ntskip) {
            std::printf("  particle %ld vel", (long)i);
            for (int d = 0; d < std::min(spec.OD, 3); ++d) std::printf(" %g", u[(size_t)d * numTargs + i]);
            std::printf("\n");
        }
    };
    auto report_err = [&](const char* name) {
        float errsum = 0.0f, errcnt = 0.0f, maxerr = 0.0f;                               // ongrav3d.cpp:782-789
        for (size_t i = 0; i < numTargs; i += ntskip) {
            const float e = u[i] - naiveu[i];
            errsum += e * e; if (e * e > maxerr) maxerr = e * e; errcnt += naiveu[i] * naiveu[i];
        }
        std::printf("error in %s (max/rms):\t%g / %g\n", name, std::sqrt(maxerr / (ntskip * errcnt / (float)numTargs)), std::sqrt(errsum / errcnt));
    };

    std::printf("\nRun the naive O(N^2) method (every %ld particles)\n", (long)ntskip);
    ALL(onb_zero_vels(ctx));
    start = now_s(); ALL(onb_naive(ctx, ntskip, &rank_flops[rk__])); double tn = now_s() - start;
    flops = rank_flops[0];                                               // (the estimate is the global formula on every rank)
    std::printf("  this run time:\t\t[%.4f] seconds\n", tn);
    std::printf("[onbody naive]:\t\t\t[%.4f] seconds\n", tn * (float)ntskip);
    std::printf("  GFlop: %.2f and GFlop/s: %.3f\n", flops * 1.e-9 * (float)ntskip, flops * 1.e-9 / tn);
    fetch(); echo();
    std::copy(u.begin(), u.begin() + numTargs, naiveu.begin());

    struct Method { const char* head; const char* tag; const char* errname; int which; float theta; int tt; };
    char h1[160], h2[200], h3[240];
    std::snprintf(h1, sizeof h1, "Run the treecode O(NlogN) with theta %g", theta1);
    std::snprintf(h2, sizeof h2, "Run the treecode O(NlogN) with %s and theta %g", withwhat.c_str(), theta2);
    std::snprintf(h3, sizeof h3, "Run the treecode O(NlogN) with %s and boxwise interactions and theta %g", withwhat.c_str(), theta3);
    const Method methods[3] = { {h1, "treecode", "treecode", 1, theta1, 1}, {h2, "treecode2", "treecode2", 2, theta2, 2}, {h3, "treecode3", "treecode3", 3, theta3, 3} };
    for (const Method& m : methods) {
        std::printf("\n%s\n", m.head);
        ALL(onb_zero_vels(ctx));
        start = now_s();
        if (m.which == 1) ALL(onb_treecode1(ctx, m.theta, &rank_flops[rk__]));
        else if (m.which == 2) ALL(onb_treecode2(ctx, m.theta, &rank_flops[rk__]));
        else ALL(onb_treecode3(ctx, m.theta, &rank_flops[rk__]));
        dt = now_s() - start;
        sum_flops();                                                     // every rank counted the lists of its own targets
        std::printf("  this run time:\t\t[%.4f] seconds\n", dt);
        std::printf("[onbody %s]:\t\t[%.4f] seconds\n", m.tag, dt);
        std::printf("  GFlop: %.3f and GFlop/s: %.3f\n", flops * 1.e-9, flops * 1.e-9 / dt);
        std::printf("[%s total]:\t\t[%.4f] seconds\n", m.tag, treetime[m.tt] + dt);
        fetch(); echo(); report_err(m.errname);
    }
    if (run_fast) {
        std::printf("\nRun the fast O(N) method with theta %g\n", theta4);
        ALL(onb_zero_vels(ctx));
        start = now_s(); ALL(onb_fastsumm(ctx, theta4)); dt = now_s() - start;
        std::printf("  this run time:\t\t[%.4f] seconds\n", dt);
        std::printf("[onbody fast]:\t\t\t[%.4f] seconds\n", dt);
        std::printf("[fast total]:\t\t\t[%.4f] seconds\n", treetime[4] + dt);
        fetch(); echo(); report_err("fastsumm");
    }
    std::printf("\nDone.\n");
    for (onb_context* q : ctxs) onb_destroy(q);
    return 0;
}
