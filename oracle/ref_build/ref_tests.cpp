/*
 * Runs the reference's OWN test suites (test/test_datatypes.h, test_expressions.h,
 * test_operators.h) on the asmjit path. TEST INFRASTRUCTURE ONLY (see ref_driver.cpp).
 * Replaces test/test.cpp, whose main() starts with a nasm pass (nasm is not installed) and
 * whose testQueries() needs tpch/datasets/sf001/lineitem.tbl (missing from the checkout).
 */
#include <iostream>
#include "operators/JitOperators.h"
#include "execute.h"
#include "test_common.h"
#include "test_datatypes.h"
#include "test_expressions.h"
#include "test_operators.h"

size_t DataBlock::Size = 2 << 20;

int main() {
    for (int threads : {1, 4, 16}) {
        testConfig.jit.emitMachineCode = true;
        testConfig.jit.numThreads = threads;
        std::cout << "== reference suites, asmjit, threads=" << threads << std::endl;
        testDatatypes();
        testExpressions();
        testOperators();
    }
    DataBlock::Size = 2 << 10;
    testConfig.jit.numThreads = 1;
    std::cout << "== reference suites, small blocks" << std::endl;
    testOperators();
    std::cout << "ALL REFERENCE SUITES PASSED" << std::endl;
    return 0;
}
