// Fateman's polynomial multiplication test, number 1: f = (1+x+y+z+t)^20, time f * (f + 1), expect 135751 terms
// (reference: benchmarks/fateman1.cpp:50-56).  Usage: fateman1 [repetitions]
#include <cstdlib>

#include "common.hpp"

int main(int argc, char **argv)
{
    using namespace pb200_bench;
    pt g;
    const pt f = fateman(20, &g);
    return run("fateman1", f, g, 135751u, argc > 1 ? std::atoi(argv[1]) : 5);
}
