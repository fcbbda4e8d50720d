// Pearce's sparse multiplication test, number 1: f = (1+x+y+2z^2+3t^3+5u^5)^12, g = (1+u+t+2z^2+3y^3+5x^5)^12,
// time f * g, expect 5821335 terms (reference: benchmarks/pearce1.cpp:51-57).
#include <cstdlib>

#include "common.hpp"

int main(int argc, char **argv)
{
    using namespace pb200_bench;
    pt g;
    const pt f = pearce(12, &g);
    return run("pearce1", f, g, 5821335u, argc > 1 ? std::atoi(argv[1]) : 3);
}
