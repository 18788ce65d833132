// Host build of pyspade_b200/csrc/gamma_fit.cuh for tests/test_gamma_fit.py: the same source the
// device kernel compiles, looped over the genes of a [S][G] table on the CPU, so that the
// restatement of scipy.stats.gamma.fit can be checked against scipy without a GPU.
#include "../pyspade_b200/csrc/gamma_fit.cuh"

extern "C" void gamma_fit_table(const double *tab, int S, long G, double sign, double *A, double *B,
                                double *C, int *status)
{
    for (long g = 0; g < G; ++g) {
        spade::GammaColumn col{tab + g, G, S, sign};
        double out[3];
        status[g] = spade::gamma_fit_column(col, out, spade::SerialReduce());
        A[g] = out[0];
        B[g] = out[1];
        C[g] = out[2];
    }
}

extern "C" double gamma_nnlf(const double *x, int n, double a, double loc, double scale)
{
    spade::GammaColumn col{x, 1, n, 1.0};
    return spade::gamma_penalized_nnlf(col, n, a, loc, scale, spade::SerialReduce());
}
