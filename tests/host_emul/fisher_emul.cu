// TEST INFRASTRUCTURE: the Fisher code of scoary_b200/csrc/fisher.cuh compiled for the HOST.
//
// fisher_two_sided() is a warp-cooperative function (four tables per warp, eight lanes each: lanes split the
// hypergeometric support, shuffles sum the lanes, a ballot drives the boundary search).  Here a "warp" is 32 host threads that
// meet at a pthread barrier inside every collective, so the very same source runs with the same
// lane-to-term assignment and the same summation order as on the GPU.  Differences to the device:
// exp() comes from the host libm (<= 1 ulp), everything else is IEEE double in the same order.
// The double-double log-factorial table is the product's own (csrc/lut.cpp, linked in).
#define SB_HOST_EMUL 1
#include <math.h>
#include <pthread.h>
#include <stdint.h>

#include <vector>

namespace {
pthread_barrier_t g_bar;
double g_d[32];
unsigned g_u[32];
thread_local int t_lane = 0;
}  // namespace

static inline double sb_emul_shfl_xor(double v, int o)
{
    g_d[t_lane] = v;
    pthread_barrier_wait(&g_bar);
    const double r = g_d[t_lane ^ o];
    pthread_barrier_wait(&g_bar);
    return r;
}
static inline double sb_emul_shfl(double v, int src)
{
    g_d[t_lane] = v;
    pthread_barrier_wait(&g_bar);
    const double r = g_d[src];
    pthread_barrier_wait(&g_bar);
    return r;
}
static inline unsigned sb_emul_ballot(bool pred)
{
    g_u[t_lane] = pred ? 1u : 0u;
    pthread_barrier_wait(&g_bar);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= g_u[l] << l;
    pthread_barrier_wait(&g_bar);
    return m;
}
static inline bool sb_emul_any(bool pred) { return sb_emul_ballot(pred) != 0u; }
// __dmul_rn / __ddiv_rn: one correctly rounded operation each, never contracted into an FMA
static inline double sb_emul_mul(double a, double b) { volatile double r = a * b; return r; }
static inline double sb_emul_div(double a, double b) { volatile double r = a / b; return r; }

#include "../../scoary_b200/csrc/fisher.cuh"

extern "C" void sb_build_logfact_dd(int32_t n, double *hi_lo);

namespace {
struct Job {
    const int32_t *tables;
    int64_t n;
    const double2 *lut;
    double *p;
    int lane;
};

void *lane_main(void *arg)      // four tables per warp, eight lanes each
{
    const Job *j = static_cast<const Job *>(arg);
    t_lane = j->lane;
    const int grp = j->lane >> 3;
    for (int64_t i = 0; i < j->n; i += 4) {
        const bool valid = i + grp < j->n;
        const int32_t *t = j->tables + 4 * (valid ? i + grp : 0);
        const double pv = sb::fisher_two_sided(j->lut, t[0], t[1], t[2], t[3], valid, j->lane);
        if (valid && (j->lane & 7) == 0) j->p[i + grp] = pv;
    }
    return nullptr;
}
}  // namespace

extern "C" {

// tables[n][4] = a, b, c, d of [[a, b], [c, d]] as the kernel passes them (tpgp, tpgn, tngp, tngn)
static int run_fisher(const int32_t *tables, int64_t n, double *p, void *(*fn)(void *));

int emul_fisher(const int32_t *tables, int64_t n, double *p) { return run_fisher(tables, n, p, lane_main); }

static int run_fisher(const int32_t *tables, int64_t n, double *p, void *(*fn)(void *))
{
    int32_t lut_n = 1;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t m = tables[4 * i] + tables[4 * i + 1] + tables[4 * i + 2] + tables[4 * i + 3];
        lut_n = m > lut_n ? m : lut_n;
    }
    std::vector<double> hl((size_t)(lut_n + 1) * 2);
    sb_build_logfact_dd(lut_n, hl.data());
    const double2 *lut = reinterpret_cast<const double2 *>(hl.data());
    pthread_barrier_init(&g_bar, nullptr, 32);
    pthread_t th[32];
    Job jobs[32];
    for (int l = 0; l < 32; ++l) {
        jobs[l] = Job{tables, n, lut, p, l};
        pthread_create(&th[l], nullptr, fn, &jobs[l]);
    }
    for (int l = 0; l < 32; ++l) pthread_join(th[l], nullptr);
    pthread_barrier_destroy(&g_bar);
    return 0;
}

}  // extern "C"
