// Fateman's test, number 2: power 30, expect 635376 terms (reference: benchmarks/fateman2.cpp:50-56).
#include <cstdlib>

#include "common.hpp"

int main(int argc, char **argv)
{
    using namespace pb200_bench;
    pt g;
    const pt f = fateman(30, &g);
    return run("fateman2", f, g, 635376u, argc > 1 ? std::atoi(argv[1]) : 3);
}
