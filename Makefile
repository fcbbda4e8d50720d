# Builds everything in-tree: the CUDA library (C ABI), the CPU oracle (test infrastructure) and the C++ host
# programs (tests of the host mirror, benchmark executables that link the C ABI).
CXX ?= g++
CXXFLAGS ?= -std=c++17 -O2 -Wall -Wextra -pthread
LIB := piranha_b200/lib/libpiranha_b200.so
HOST_HDRS := $(wildcard piranha_b200/host/*.hpp) include/piranha_b200.h
BIN := build/bin
PROGS := $(BIN)/host_logic_test $(BIN)/multiplier_test $(BIN)/fateman1 $(BIN)/fateman2 $(BIN)/pearce1 $(BIN)/fateman1_rational $(BIN)/pearce1_rational $(BIN)/fateman1_dynamic
LINK := -Lpiranha_b200/lib -lpiranha_b200 -Wl,-rpath,'$$ORIGIN/../../piranha_b200/lib'

all: $(LIB) oracle host

$(LIB): $(wildcard piranha_b200/csrc/*) include/piranha_b200.h piranha_b200/host/kronecker_array.hpp
	./build.sh

oracle:
	$(MAKE) -C oracle -s

host: $(PROGS)

$(BIN)/%_test: tests/cpp/%_test.cpp tests/cpp/check.hpp $(HOST_HDRS) $(LIB)
	mkdir -p $(BIN)
	$(CXX) $(CXXFLAGS) -Iinclude -o $@ $< $(LINK)

$(BIN)/%: benchmarks/%.cpp benchmarks/common.hpp $(HOST_HDRS) $(LIB)
	mkdir -p $(BIN)
	$(CXX) $(CXXFLAGS) -Iinclude -o $@ $< $(LINK)

clean:
	rm -rf build piranha_b200/lib oracle/_build

.PHONY: all oracle host clean
