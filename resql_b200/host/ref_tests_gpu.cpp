/*
 * resql-reftests-gpu: the reference's OWN test suites (test/test_datatypes.h, test_expressions.h,
 * test_operators.h, compiled from the reference checkout where it lies) with every plan execution
 * routed through the drop-in: executeSelectPlan -> executeSelectPlanGpu (gpu_executor.h) -> C ABI
 * -> sm_100a kernels. This is the literal drop-in test: the suites build RelOperator trees by hand,
 * call executeSelectAndCheckRelation (test_common.h:222-230) and compare relations with the
 * reference's own comparison code.
 *
 * Scalar expression tests (test_common.h:66-91 compile one expression with the Flounder JIT) run
 * as a leaf projection (projection.h:49-58) through the same GPU path.
 */
#include <iostream>
#include "operators/JitOperators.h"
#include "execute.h"
#include "gpu_executor.h"

static std::unique_ptr<SelectResult> executeSelectPlanOnGpu ( RelOperator* root, bool requestAll, Database& db,
                                                             DBConfig config = DBConfig() ) {
    return executeSelectPlanGpu ( root, requestAll, db, config );
}

/* route the suites' calls; their CPU scalar-expression helper keeps its name with a suffix */
#define executeSelectPlan executeSelectPlanOnGpu
#define executeAndCheckExpression executeAndCheckExpressionFlounder
#include "test_common.h"
#undef executeAndCheckExpression

void executeAndCheckExpression ( std::string name, Expr* expr, std::string reference ) {
    Database db;
    ExprVec select = { expr };
    RelOperator* root = new MaterializeOp ( new ProjectionOp ( select, nullptr ) );
    std::unique_ptr<SelectResult> res = executeSelectPlanGpu ( root, true, db, testConfig );
    Relation& rel = *res->relation;
    if ( rel.tupleNum() != 1 || rel._schema._attribs.size() != 1 ) {
        std::cout << name << ": leaf projection returned " << rel.tupleNum() << " tuples" << std::endl;
        fail_test();
    }
    Relation::ReadIterator readIt ( &rel );
    Data* t = readIt.get();
    auto atts = AttributeIterator::getAll ( rel._schema );
    checkSerialized ( name, atts[0].serialize ( t ), reference );
}

#include "test_datatypes.h"
#include "test_expressions.h"
#include "test_operators.h"

size_t DataBlock::Size = 2 << 20;

int main() {
    std::cout << "== reference suites through executeSelectPlanGpu" << std::endl;
    testDatatypes();
    testExpressions();
    testOperators();
    DataBlock::Size = 2 << 10;
    std::cout << "== reference suites through executeSelectPlanGpu, small blocks" << std::endl;
    testOperators();
    rq_shutdown();
    std::cout << "ALL REFERENCE SUITES PASSED ON THE GPU PATH" << std::endl;
    return 0;
}
