/**
 * @file
 * Host shim: drop-in replacement for the Flounder/asmjit span of the reference's
 * executeSelectPlan (src/execute.h:213-247) that runs the plan on a B200 through the C ABI in
 * include/resql_b200.h.
 *
 * This header is OUR code, written against the reference's public operator API. It must be
 * included in the reference's single translation unit AFTER "operators/JitOperators.h" and
 * "execute.h" (the reference is header-only and can only live in one TU). It never includes or
 * touches src/flounder/, JitContextFlounder::compile/execute or qlib/hash.h / qlib/sort.h.
 *
 *   std::unique_ptr<SelectResult> executeSelectPlanGpu(RelOperator* root, bool requestAll,
 *                                                      Database& db, DBConfig config);
 *
 * has the signature, ownership rules and error behaviour of executeSelectPlan: the callee owns the
 * plan and frees it with root->deletePlan() on success and on error (execute.h:237,242), errors
 * are thrown as ResqlError by value, the result is a Relation in the reference's row format
 * (Schema with stringsByVal=true) so printRelation / tofile / the reference tests work unchanged.
 *
 * What it does:
 *   1. steps 1-3 of executeSelectPlan unchanged: defineExpressionsForPlan, deriveExpressionTypes,
 *      unifyExpressions (execute.h:222-227) - so the lowering reads the same typed Expr trees,
 *      including the TYPECAST nodes the type system inserted;
 *   2. walks the operator tree in the reference's produce/consume order (scan.h:227,
 *      selection.h:37-70, projection.h:40-72, aggregation.h:155-343, hashjoin.h:97-279,
 *      materialize.h:67-220, orderby.h:96-136), recomputing the attribute requests and operator
 *      schemas that the reference derives inside Flounder codegen, handing out expression ids in
 *      the same order (addExpressionIds, expressions.h:1354) so output column names match,
 *      and emits one rq_pipeline per reference pipeline with a typed postfix program per Expr;
 *   3. mirrors every scanned Relation into device columns once (rq_table_upload_rows), keyed by
 *      (Relation*, tuple count);
 *   4. calls rq_plan_execute and re-packs rq_result into a Relation.
 *
 * Environment switches (test tooling):
 *   RESQL_B200_DUMP_PLAN=<file>  write the lowered plan as JSON (fixtures in tests/golden/plans)
 *   RESQL_B200_DRY=1             lower (and dump) only; return an empty relation, no GPU needed
 */
#pragma once

#include <map>
#include <set>
#include <string>
#include <vector>
#include <sstream>
#include <fstream>
#include <cstdlib>
#include <cstring>

#include <unistd.h>
#include "resql_b200.h"

namespace rqshim {

struct ValueInfo {
    int     node;
    SqlType type;
};

struct PipelineDesc {
    int source_kind = 0, source_id = 0, source_id2 = 0;
    std::vector<rq_node>  nodes;
    std::vector<int32_t>  args;
    int sink_kind = 0;
    std::vector<rq_value> keys, vals;
    int64_t size_hint = 0;
};

struct TableDesc {
    std::string name;                 /* relation name as in Database::relations */
    Relation*   rel;
    std::vector<Attribute> attrs;     /* uploaded columns, table order */
};

static int sqlTag ( SqlType t ) { return (int) t.tag; }  /* RQ_SQL_* uses the SqlType::Tag order */

static int sqlWidth ( SqlType t ) {
    switch ( t.tag ) {
        case SqlType::CHAR:    return (int) t.charSpec().num;
        case SqlType::VARCHAR: return (int) t.varcharSpec().num;
        case SqlType::DECIMAL: return ( (int) t.decimalSpec().precision << 8 ) | (int) t.decimalSpec().scale;
        default: return 0;
    }
}

static bool isStringType ( SqlType t ) {
    return t.tag == SqlType::VARCHAR || ( t.tag == SqlType::CHAR && t.charSpec().num > 1 );
}


class Lowering {
public:
    Database&   db;
    bool        requestAll;

    std::vector<TableDesc>     tables;
    std::vector<PipelineDesc>  pipelines;
    std::string                strpool;
    std::vector<rq_order_key>  order;
    int64_t                    limit = -1;
    Schema                     resultSchema;

    /* state of the pipeline under construction */
    int cur = -1;
    std::map<std::string, ValueInfo> env;       /* symbol table: name -> value (ctx.symbolTable) */
    std::map<RelOperator*, Schema> schemas;     /* operator output schemas (op->_schema)         */
    std::map<RelOperator*, SymbolSet> requests;
    std::map<RelOperator*, int> joinCalls;
    std::map<RelOperator*, int> joinBuildPipe;
    std::map<RelOperator*, int> matPipe;        /* inner MaterializeOp -> pipeline that fills it */
    int exprIdGen = 1;                          /* RelationalContext::exprIdGen                  */

    Lowering ( Database& db, bool requestAll ) : db ( db ), requestAll ( requestAll ) {}

    /* ---- program emission ------------------------------------------------------------- */
    PipelineDesc& P() { return pipelines[cur]; }

    int node ( int op, int a = 0, int b = 0, int c = 0, int64_t imm = 0 ) {
        rq_node n; n.op = op; n.a = a; n.b = b; n.c = c; n.imm = imm;
        P().nodes.push_back ( n );
        return (int) P().nodes.size() - 1;
    }

    int constNode ( int64_t v ) { return node ( RQ_OP_CONST, 0, 0, 0, v ); }

    int strNode ( const char* s ) {
        int64_t off = (int64_t) strpool.size();
        strpool.append ( s );
        strpool.push_back ( '\0' );
        return node ( RQ_OP_CONST_STR, 0, 0, 0, off );
    }

    void addIds ( Expr* e ) { if ( e->id == 0 ) e->id = ++exprIdGen; }   /* expressions.h:1354 */

    int newPipeline ( int sourceKind, int sourceId ) {
        pipelines.emplace_back();
        cur = (int) pipelines.size() - 1;
        P().source_kind = sourceKind;
        P().source_id = sourceId;
        env.clear();
        return cur;
    }

    /* emitConstant (ExpressionsJitFlounder.h:248-291) */
    int lowerConstant ( Expr* e ) {
        switch ( e->type.tag ) {
            case SqlType::DECIMAL: return constNode ( e->value.decimalData );
            case SqlType::DATE:    return constNode ( (int64_t) (int32_t) e->value.dateData );
            case SqlType::BOOL:    return constNode ( e->value.boolData );
            case SqlType::BIGINT:  return constNode ( e->value.bigintData );
            case SqlType::INT:     return constNode ( e->value.intData );
            case SqlType::CHAR:
                if ( e->type.charSpec().num > 1 ) return strNode ( e->value.charData );
                return constNode ( (unsigned char) e->value.charData[0] );
            case SqlType::VARCHAR: return strNode ( e->value.varcharData );
            default:
                throw ResqlError ( "Constant code generation not implemented for datatype" );
        }
    }

    static bool isArith ( SqlType t ) { return t.tag == SqlType::DECIMAL || t.tag == SqlType::BIGINT; }
    static bool isOrdered ( SqlType t ) {
        return t.tag == SqlType::DECIMAL || t.tag == SqlType::DATE || t.tag == SqlType::BIGINT;
    }

    /* emitTypecast (ExpressionsJitFlounder.h:760-882) */
    int lowerTypecast ( SqlType from, SqlType to, int child ) {
        static const int64_t f[] = { 1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000 };
        if ( to.tag == SqlType::DECIMAL ) {
            if ( from.tag == SqlType::DECIMAL ) {
                int fs = from.decimalSpec().scale, ts = to.decimalSpec().scale;
                if ( ts == fs ) return child;
                if ( ts >= fs ) {
                    if ( ts - fs > 8 ) throw ResqlError ( "decimal typecast beyond 10^8" );
                    return node ( RQ_OP_MUL, child, constNode ( f [ ts - fs ] ) );
                }
                if ( fs - ts > 8 ) throw ResqlError ( "decimal typecast beyond 10^8" );
                return node ( RQ_OP_DIV, child, constNode ( f [ fs - ts ] ) );
            }
            if ( from.tag == SqlType::BIGINT ) {
                return node ( RQ_OP_MUL, child, constNode ( f [ to.decimalSpec().scale ] ) );
            }
            throw ResqlError ( "emitTypecastToDECIMAL(..) code generation not implemented for datatype" );
        }
        if ( to.tag == SqlType::BIGINT ) {
            if ( from.tag == SqlType::INT ) return child;      /* movsxd: values are already sign-extended */
            if ( from.tag == SqlType::DECIMAL ) return node ( RQ_OP_DIV, child, constNode ( f [ from.decimalSpec().scale ] ) );
            if ( from.tag == SqlType::BIGINT ) return child;
            throw ResqlError ( "emitTypecastToBIGINT(..) code generation not implemented for datatype" );
        }
        throw ResqlError ( "emitTypecast(..) code generation not implemented for datatype" );
    }

    /* emitExpression (ExpressionsJitFlounder.h:1080-1114) */
    int lower ( Expr* e ) {
        if ( e->type.tag == SqlType::NT && e->tag != Expr::STAR ) {
            throw ResqlError ( "Expression type undefined. Have you derived the expression types?" );
        }
        std::string name = getExpressionName ( e );
        auto it = env.find ( name );
        if ( it != env.end() ) return it->second.node;

        switch ( e->structureTag ) {
        case Expr::LITERAL:
            if ( e->tag == Expr::ATTRIBUTE ) throw ResqlError ( "Attribute " + e->symbol + " not available in pipeline" );
            if ( e->tag == Expr::CONSTANT ) return lowerConstant ( e );
            throw ResqlError ( std::string ( "emitExpressionLiteral(..) not implemented for expression type" ) + exprTagNames [ e->tag ] );

        case Expr::UNARY: {
            if ( e->tag == Expr::COUNT ) return constNode ( 1 );             /* emitCount :710 */
            int child = lower ( e->child );
            switch ( e->tag ) {
                case Expr::SUM: case Expr::AVG: case Expr::MIN: case Expr::MAX: case Expr::AS:
                    return child;
                case Expr::TYPECAST:
                    return lowerTypecast ( e->child->type, e->type, child );
                default:
                    throw ResqlError ( std::string ( "emitExpression(..) not implemented for expression type" ) + exprTagNames [ e->tag ] );
            }
        }

        case Expr::BINARY: {
            int l = lower ( e->child );
            int r = lower ( e->child->next );
            SqlType type = e->type;
            SqlType opType = e->child->type;
            switch ( e->tag ) {
                case Expr::ADD: if ( !isArith ( type ) ) throw ResqlError ( "ADD code generation not implemented for datatype" ); return node ( RQ_OP_ADD, l, r );
                case Expr::SUB: if ( !isArith ( type ) ) throw ResqlError ( "SUB code generation not implemented for datatype" ); return node ( RQ_OP_SUB, l, r );
                case Expr::MUL: if ( !isArith ( type ) ) throw ResqlError ( "MUL code generation not implemented for datatype" ); return node ( RQ_OP_MUL, l, r );
                case Expr::DIV: if ( !isArith ( type ) ) throw ResqlError ( "DIV code generation not implemented for datatype" ); return node ( RQ_OP_DIV, l, r );
                case Expr::AND: return node ( RQ_OP_AND, l, r );
                case Expr::OR:  return node ( RQ_OP_OR, l, r );
                case Expr::LT:  if ( !isOrdered ( opType ) ) throw ResqlError ( "LESS_THAN code generation not implemented for datatype" ); return node ( RQ_OP_LT, l, r );
                case Expr::LE:  if ( !isOrdered ( opType ) ) throw ResqlError ( "LESS_THAN_OR_EQUAL code generation not implemented for datatype" ); return node ( RQ_OP_LE, l, r );
                case Expr::GT:  if ( !isOrdered ( opType ) ) throw ResqlError ( "GREATER_THAN code generation not implemented for datatype" ); return node ( RQ_OP_GT, l, r );
                case Expr::GE:  if ( !isOrdered ( opType ) ) throw ResqlError ( "GREATER_THAN_OR_EQUAL code generation not implemented for datatype" ); return node ( RQ_OP_GE, l, r );
                case Expr::EQ:
                case Expr::NEQ: {
                    bool neq = e->tag == Expr::NEQ;
                    switch ( opType.tag ) {
                        case SqlType::DECIMAL: case SqlType::INT: case SqlType::BIGINT:
                        case SqlType::BOOL: case SqlType::DATE:
                            return node ( neq ? RQ_OP_NEQ : RQ_OP_EQ, l, r );
                        case SqlType::CHAR:
                            if ( opType.charSpec().num > 1 ) return node ( neq ? RQ_OP_NEQ_CHAR : RQ_OP_EQ_CHAR, l, r );
                            return node ( neq ? RQ_OP_NEQ : RQ_OP_EQ, l, r );
                        case SqlType::VARCHAR:
                            return node ( neq ? RQ_OP_NEQ_VARCHAR : RQ_OP_EQ_VARCHAR, l, r );
                        default:
                            throw ResqlError ( "EQUALS code generation not implemented for datatype" );
                    }
                }
                case Expr::LIKE: return node ( RQ_OP_LIKE, l, r );
                default:
                    throw ResqlError ( std::string ( "emitExpressionBinary(..) not implemented for expression type" ) + exprTagNames [ e->tag ] );
            }
        }

        case Expr::OTHER: {
            if ( e->tag != Expr::CASE ) throw ResqlError ( "emitExpressionOther(..) not implemented for expression type" );
            /* emitCase (:720-754): WHEN/THEN arms tested in order; without ELSE the reference leaves
             * the result register undefined - this engine defines it as 0 */
            std::vector<std::pair<int,int>> arms;
            Expr* child = e->child;
            while ( child != nullptr && child->tag == Expr::WHENTHEN ) {
                int w = lower ( child->child );
                int t = lower ( child->child->next );
                arms.push_back ( { w, t } );
                child = child->next;
            }
            int res = ( child != nullptr ) ? lower ( child ) : constNode ( 0 );
            for ( int i = (int) arms.size() - 1; i >= 0; i-- ) {
                res = node ( RQ_OP_SELECT, arms[i].first, arms[i].second, res );
            }
            return res;
        }
        default:
            throw ResqlError ( "emitExpression(..) not implemented" );
        }
    }

    /* evalExpressions (ValuesJitFlounder.h:471-482) */
    std::vector<std::pair<std::string, ValueInfo>> evalExpressions ( std::vector<Expr*>& exprs ) {
        std::vector<std::pair<std::string, ValueInfo>> res;
        for ( Expr* e : exprs ) {
            addIds ( e );
            int n = lower ( e );
            res.push_back ( { getExpressionName ( e ), { n, e->type } } );
        }
        return res;
    }

    static Schema schemaOf ( std::vector<std::pair<std::string, ValueInfo>>& vals ) {
        std::vector<Attribute> atts;
        for ( auto& v : vals ) atts.push_back ( { v.first, v.second.type } );
        return Schema ( atts, true );
    }

    rq_value valueOf ( int nodeIdx, SqlType t, int kind = 0 ) {
        rq_value v; v.node = nodeIdx; v.kind = kind; v.sql_type = sqlTag ( t ); v.width = sqlWidth ( t );
        return v;
    }

    /* ---- produce / consume ---------------------------------------------------------------- */
    int tableIndex ( ScanOp* scan, std::vector<Attribute>& attrs ) {
        std::string relName;
        for ( auto& r : db.relations ) if ( &r.second == scan->_rel ) relName = r.first;
        if ( relName.empty() ) relName = scan->relationName.empty() ? ( "rel" + std::to_string ( tables.size() ) ) : scan->relationName;
        tables.push_back ( { relName, scan->_rel, attrs } );
        return (int) tables.size() - 1;
    }

    void produce ( RelOperator* op, SymbolSet request ) {
        if ( auto* scan = dynamic_cast<ScanOp*> ( op ) ) {
            if ( requestAll ) request = {};
            /* Values::dematerialize with a request: requested attributes in table order;      *
             * an empty request means all attributes (ValuesJitFlounder.h:396)                 */
            std::vector<Attribute> attrs;
            for ( auto& a : scan->_rel->_schema._attribs ) {
                if ( request.empty() || request.count ( a.name ) ) attrs.push_back ( a );
            }
            int t = tableIndex ( scan, attrs );
            newPipeline ( RQ_SRC_TABLE, t );
            for ( size_t i = 0; i < attrs.size(); i++ ) {
                env [ attrs[i].name ] = { node ( RQ_OP_COL, (int) i ), attrs[i].type };
            }
            schemas [ op ] = Schema ( attrs, true );
            consume ( op->_parent, op );
            return;
        }
        if ( auto* sel = dynamic_cast<SelectionOp*> ( op ) ) {
            requests [ op ] = request;
            SymbolSet selReq = extractRequiredAttributes ( sel->_condition );
            produce ( sel->_child, symbolSetUnion ( request, selReq ) );
            return;
        }
        if ( auto* proj = dynamic_cast<ProjectionOp*> ( op ) ) {
            if ( proj->_child == nullptr ) {
                /* leaf projection (projection.h:49-58): "creates a single-line table from constants" */
                newPipeline ( RQ_SRC_ONE_ROW, 0 );
                consume ( op, nullptr );
                return;
            }
            SymbolSet req = extractRequiredAttributes ( proj->_expr );
            produce ( proj->_child, req );
            return;
        }
        if ( auto* agg = dynamic_cast<AggregationOp*> ( op ) ) {
            SymbolSet aggReq = extractRequiredAttributes ( agg->_aggExpr );
            SymbolSet groupReq = extractRequiredAttributes ( agg->_groupExpr );
            produce ( agg->_child, symbolSetUnion ( aggReq, groupReq ) );
            consumeAggregate ( agg );
            return;
        }
        if ( auto* hj = dynamic_cast<HashJoinOp*> ( op ) ) {
            requests [ op ] = request;
            SymbolSet joinReq = extractRequiredAttributes ( hj->_equalities );
            SymbolSet allReq = symbolSetUnion ( request, joinReq );
            joinCalls [ op ] = 0;
            produce ( hj->_lChild, allReq );
            produce ( hj->_rChild, allReq );
            return;
        }
        if ( auto* nlj = dynamic_cast<NestedLoopsJoinOp*> ( op ) ) {
            /* nestedloopsjoin.h:53-67: both (materialized) children are produced with the parent's request */
            joinCalls [ op ] = 0;
            produce ( nlj->_lChild, request );
            produce ( nlj->_rChild, request );
            return;
        }
        if ( dynamic_cast<OrderByOp*> ( op ) ) {
            auto* ob = static_cast<OrderByOp*> ( op );
            produce ( ob->_child, request );
            Schema& s = schemas [ ob->_child ];
            schemas [ op ] = s;
            for ( Expr* e : ob->_orderExpressions ) {
                if ( !s.contains ( e->child->symbol ) ) throw ResqlError ( "Order By attribute not found." );
                rq_order_key k;
                k.column = -1;
                for ( size_t i = 0; i < s._attribs.size(); i++ ) if ( s._attribs[i].name == e->child->symbol ) { k.column = (int) i; break; }
                k.ascending = e->tag != Expr::DESC;
                order.push_back ( k );
            }
            if ( ob->_hasLimitClause ) limit = (int64_t) ob->_limit;
            return;
        }
        if ( op->isMaterializedOperator() ) {   /* MaterializeOp (its tag is SELECTION, materialize.h:33) */
            produce ( op->_child, request );
            return;
        }
        throw ResqlError ( "GPU engine: operator " + op->name() + " is not on the hot path (SURVEY section 8f)" );
    }

    void consume ( RelOperator* op, RelOperator* from ) {
        if ( op == nullptr ) throw ResqlError ( "GPU engine: plan root must be a materializing operator" );
        if ( auto* sel = dynamic_cast<SelectionOp*> ( op ) ) {
            Schema s = schemas [ sel->_child ];
            if ( !requestAll ) s = s.prune ( requests [ op ] );
            schemas [ op ] = s;
            addIds ( sel->_condition );
            int c = lower ( sel->_condition );
            node ( RQ_OP_FILTER, c );
            consume ( op->_parent, op );
            return;
        }
        if ( auto* proj = dynamic_cast<ProjectionOp*> ( op ) ) {
            auto vals = evalExpressions ( proj->_expr );
            for ( auto& v : vals ) env [ v.first ] = v.second;
            schemas [ op ] = schemaOf ( vals );
            consume ( op->_parent, op );
            return;
        }
        if ( auto* agg = dynamic_cast<AggregationOp*> ( op ) ) {
            auto groupVals = evalExpressions ( agg->_groupExpr );
            auto aggVals = evalExpressions ( agg->_splitAggExpr );
            PipelineDesc& p = P();
            p.sink_kind = RQ_SINK_AGG;
            p.size_hint = (int64_t) agg->getSize();
            for ( auto& g : groupVals ) p.keys.push_back ( valueOf ( g.second.node, g.second.type ) );
            for ( size_t i = 0; i < aggVals.size(); i++ ) {
                Expr* e = agg->_splitAggExpr[i];
                int kind;
                switch ( e->tag ) {
                    case Expr::SUM:   kind = RQ_AGG_SUM; break;
                    case Expr::COUNT: kind = RQ_AGG_COUNT; break;
                    case Expr::MIN:   kind = RQ_AGG_MIN; break;
                    case Expr::MAX:   kind = RQ_AGG_MAX; break;
                    default: throw ResqlError ( "Aggregation type not implemented in updateAggregates(..)." );
                }
                if ( kind == RQ_AGG_SUM && !isArith ( aggVals[i].second.type ) ) throw ResqlError ( "ADD code generation not implemented for datatype" );
                if ( ( kind == RQ_AGG_MIN || kind == RQ_AGG_MAX ) && !isOrdered ( aggVals[i].second.type ) ) throw ResqlError ( "LESS_THAN code generation not implemented for datatype" );
                p.vals.push_back ( valueOf ( aggVals[i].second.node, aggVals[i].second.type, kind ) );
            }
            aggGroupVals [ agg ] = groupVals;
            aggAggVals [ agg ] = aggVals;
            aggPipe [ agg ] = cur;
            return;    /* pipeline ends here; consumeAggregate continues after produce(child) returns */
        }
        if ( auto* hj = dynamic_cast<HashJoinOp*> ( op ) ) {
            int call = ++joinCalls [ op ];
            if ( call == 1 ) {
                /* build (hashjoin.h:226-256): keys = left sides, payload = all left-child attributes */
                auto left = equalitiesLeftSide ( hj->_equalities );
                auto buildKeys = evalExpressions ( left );
                PipelineDesc& p = P();
                p.sink_kind = RQ_SINK_BUILD;
                p.size_hint = (int64_t) hj->_lChild->getSize();
                for ( auto& k : buildKeys ) p.keys.push_back ( valueOf ( k.second.node, k.second.type ) );
                Schema& ls = schemas [ hj->_lChild ];
                for ( auto& a : ls._attribs ) {
                    auto it = env.find ( a.name );
                    if ( it == env.end() ) throw ResqlError ( "GPU engine: build attribute " + a.name + " not available" );
                    p.vals.push_back ( valueOf ( it->second.node, it->second.type ) );
                }
                joinBuildPipe [ op ] = cur;
                return;
            }
            if ( call == 2 ) {
                /* probe (hashjoin.h:258-279) */
                Schema ls = schemas [ hj->_lChild ];
                Schema s = ls.join ( schemas [ hj->_rChild ] );
                if ( !requestAll ) s = s.prune ( requests [ op ] );
                schemas [ op ] = s;
                auto right = equalitiesRightSide ( hj->_equalities );
                auto probeKeys = evalExpressions ( right );
                PipelineDesc& p = P();
                int first = (int) p.args.size();
                for ( auto& k : probeKeys ) p.args.push_back ( k.second.node );
                int pn = node ( RQ_OP_PROBE, joinBuildPipe [ op ], first, (int) probeKeys.size(), hj->_singleMatch ? 1 : 0 );
                for ( size_t i = 0; i < ls._attribs.size(); i++ ) {
                    env [ ls._attribs[i].name ] = { node ( RQ_OP_PAYLOAD, pn, (int) i ), ls._attribs[i].type };
                }
                consume ( op->_parent, op );
                return;
            }
            throw ResqlError ( "HashJoin::consumeFlounder(..) called more than 2 times." );
        }
        if ( auto* nlj = dynamic_cast<NestedLoopsJoinOp*> ( op ) ) {
            /* nestedloopsjoin.h:70-93: after both children are materialized every pair of tuples is
             * produced (schema = left ++ right), the optional condition drops pairs */
            int call = ++joinCalls [ op ];
            if ( call < 2 ) return;
            Schema ls = schemas [ nlj->_lChild ], rs = schemas [ nlj->_rChild ];
            newPipeline ( RQ_SRC_CROSS, matPipe [ nlj->_lChild ] );
            P().source_id2 = matPipe [ nlj->_rChild ];
            int col = 0;
            for ( auto& a : ls._attribs ) env [ a.name ] = { node ( RQ_OP_COL, col++ ), a.type };
            for ( auto& a : rs._attribs ) env [ a.name ] = { node ( RQ_OP_COL, col++ ), a.type };
            schemas [ op ] = ls.join ( rs );
            if ( nlj->_condition != nullptr ) {
                addIds ( nlj->_condition );
                node ( RQ_OP_FILTER, lower ( nlj->_condition ) );
            }
            consume ( op->_parent, op );
            return;
        }
        if ( dynamic_cast<OrderByOp*> ( op ) ) return;
        if ( op->isMaterializedOperator() && op->_parent != nullptr && dynamic_cast<NestedLoopsJoinOp*> ( op->_parent ) ) {
            /* MaterializeOp wrapped around a nested-loops join child (nestedloopsjoin.h:25-26) */
            Schema s = schemas [ op->_child ];
            schemas [ op ] = s;
            PipelineDesc& p = P();
            p.sink_kind = RQ_SINK_MATERIALIZE;
            for ( auto& a : s._attribs ) {
                auto it = env.find ( a.name );
                if ( it == env.end() ) throw ResqlError ( "GPU engine: attribute " + a.name + " not available" );
                p.vals.push_back ( valueOf ( it->second.node, a.type ) );
            }
            matPipe [ op ] = cur;
            consume ( op->_parent, op );
            return;
        }
        if ( op->isMaterializedOperator() ) {
            Schema s = schemas [ op->_child ];
            schemas [ op ] = s;
            PipelineDesc& p = P();
            p.sink_kind = RQ_SINK_MATERIALIZE;
            for ( auto& a : s._attribs ) {
                auto it = env.find ( a.name );
                if ( it == env.end() ) throw ResqlError ( "GPU engine: result attribute " + a.name + " not available" );
                p.vals.push_back ( valueOf ( it->second.node, a.type ) );
            }
            resultSchema = Schema ( s._attribs, true );
            auto* mat = static_cast<MaterializeOp*> ( op );
            if ( mat->_hasLimitClause ) limit = (int64_t) mat->_limit;
            return;
        }
        throw ResqlError ( "GPU engine: operator " + op->name() + " is not on the hot path (SURVEY section 8f)" );
    }

    std::map<RelOperator*, std::vector<std::pair<std::string, ValueInfo>>> aggGroupVals, aggAggVals;
    std::map<RelOperator*, int> aggPipe;

    /* consumeAggregateFlounder + mergeAverages (aggregation.h:298-343, :207-238) */
    void consumeAggregate ( AggregationOp* agg ) {
        auto& groupVals = aggGroupVals [ agg ];
        auto& aggVals = aggAggVals [ agg ];
        int src = aggPipe [ agg ];
        newPipeline ( RQ_SRC_PIPELINE, src );
        std::vector<std::pair<std::string, ValueInfo>> out;
        int col = 0;
        for ( auto& g : groupVals ) {
            out.push_back ( { g.first, { node ( RQ_OP_COL, col++ ), g.second.type } } );
        }
        size_t i = 0;
        for ( Expr* e : agg->_aggExpr ) {
            if ( e->tag == Expr::AVG ) {
                addIds ( e );
                int sum = node ( RQ_OP_COL, col++ );
                int count = node ( RQ_OP_COL, col++ );
                SqlType st = aggVals[i].second.type;
                if ( st.tag != SqlType::BIGINT && st.tag != SqlType::DECIMAL ) throw ResqlError ( "getAvgFromSumAndCount(..) not supported for datatype" );
                int scaled = node ( RQ_OP_MUL, sum, constNode ( 100 ) );
                int avg = node ( RQ_OP_DIV, scaled, count );
                out.push_back ( { getExpressionName ( e ), { avg, e->type } } );
                i += 2;
            }
            else {
                out.push_back ( { aggVals[i].first, { node ( RQ_OP_COL, col++ ), aggVals[i].second.type } } );
                i += 1;
            }
        }
        for ( auto& v : out ) env [ v.first ] = v.second;
        schemas [ agg ] = schemaOf ( out );
        consume ( agg->_parent, agg );
    }

    /* ---- JSON dump (plan fixtures) ------------------------------------------------------------ */
    static std::string jsonEscape ( const std::string& s ) {
        std::ostringstream o;
        for ( unsigned char c : s ) {
            if ( c == '"' || c == '\\' ) o << '\\' << c;
            else if ( c < 0x20 || c >= 0x7f ) { char b[8]; snprintf ( b, sizeof b, "\\u%04x", c ); o << b; }
            else o << c;
        }
        return o.str();
    }

    std::string toJson ( ) {
        std::ostringstream o;
        o << "{\n \"tables\": [";
        for ( size_t t = 0; t < tables.size(); t++ ) {
            o << ( t ? ", " : "" ) << "{\"name\": \"" << tables[t].name << "\", \"columns\": [";
            for ( size_t i = 0; i < tables[t].attrs.size(); i++ ) o << ( i ? ", " : "" ) << "\"" << tables[t].attrs[i].name << "\"";
            o << "]}";
        }
        o << "],\n \"pipelines\": [\n";
        for ( size_t p = 0; p < pipelines.size(); p++ ) {
            PipelineDesc& pd = pipelines[p];
            o << "  {\"source_kind\": " << pd.source_kind << ", \"source_id\": " << pd.source_id << ", \"source_id2\": " << pd.source_id2 << ", \"sink_kind\": " << pd.sink_kind
              << ", \"size_hint\": " << pd.size_hint << ",\n   \"nodes\": [";
            for ( size_t i = 0; i < pd.nodes.size(); i++ ) {
                rq_node& n = pd.nodes[i];
                o << ( i ? ", " : "" ) << "[" << n.op << ", " << n.a << ", " << n.b << ", " << n.c << ", " << n.imm << "]";
            }
            o << "],\n   \"args\": [";
            for ( size_t i = 0; i < pd.args.size(); i++ ) o << ( i ? ", " : "" ) << pd.args[i];
            o << "],\n   \"keys\": [";
            for ( size_t i = 0; i < pd.keys.size(); i++ ) o << ( i ? ", " : "" ) << "[" << pd.keys[i].node << ", " << pd.keys[i].kind << ", " << pd.keys[i].sql_type << ", " << pd.keys[i].width << "]";
            o << "],\n   \"vals\": [";
            for ( size_t i = 0; i < pd.vals.size(); i++ ) o << ( i ? ", " : "" ) << "[" << pd.vals[i].node << ", " << pd.vals[i].kind << ", " << pd.vals[i].sql_type << ", " << pd.vals[i].width << "]";
            o << "]}" << ( p + 1 < pipelines.size() ? "," : "" ) << "\n";
        }
        o << " ],\n \"order\": [";
        for ( size_t i = 0; i < order.size(); i++ ) o << ( i ? ", " : "" ) << "[" << order[i].column << ", " << order[i].ascending << "]";
        o << "],\n \"limit\": " << limit << ",\n \"strpool\": \"" << jsonEscape ( strpool ) << "\",\n \"result_names\": [";
        for ( size_t i = 0; i < resultSchema._attribs.size(); i++ ) o << ( i ? ", " : "" ) << "\"" << resultSchema._attribs[i].name << "\"";
        o << "],\n \"result_types\": [";
        for ( size_t i = 0; i < resultSchema._attribs.size(); i++ ) o << ( i ? ", " : "" ) << "\"" << serializeType ( resultSchema._attribs[i].type ) << "\"";
        o << "]\n}\n";
        return o.str();
    }
};


/* ---- device mirror of the row store (dbdata.h) ----------------------------------------------- */
struct MirrorKey {
    Relation* rel; size_t tuples; std::string cols; bool shard;
    bool operator< ( const MirrorKey& o ) const {
        if ( rel != o.rel ) return rel < o.rel;
        if ( tuples != o.tuples ) return tuples < o.tuples;
        if ( shard != o.shard ) return shard < o.shard;
        return cols < o.cols;
    }
};
static std::map<MirrorKey, rq_table*>& mirrors() { static std::map<MirrorKey, rq_table*> m; return m; }
/* Multi-GPU (`gpus=N`, the counterpart of the reference's `threads=N` control variable,
 * execute.h:454-474): one process per GPU, every process runs the same statements on the same row
 * store. For a select the table scanned by the last aggregation's pipeline (the fact table) is
 * mirrored as THIS rank's row range, all other tables completely, and the plan runs with
 * RQ_PLAN_SHARDED: partial results are merged over NCCL inside the library and every rank holds
 * the full result. rank / world / the file that carries the NCCL id are set by the host program
 * (resql_b200_driver.cpp) before the first select. */
struct Group { int rank = 0, world = 1; std::string idFile; bool joined = false; };
static Group& group() { static Group g; return g; }
static bool& engineUp() { static bool up = false; return up; }
static double& lastLoadMs() { static double ms = 0; return ms; }

static void check ( int rc ) {
    if ( rc != RQ_OK ) throw ResqlError ( std::string ( "GPU engine: " ) + rq_last_error() );
}

static rq_table* mirrorTable ( TableDesc& t, bool shard = false ) {
    std::string cols;
    for ( auto& a : t.attrs ) cols += a.name + ",";
    MirrorKey key { t.rel, t.rel->tupleNum(), cols, shard };
    auto it = mirrors().find ( key );
    if ( it != mirrors().end() ) return it->second;
    Schema& s = t.rel->_schema;
    std::vector<int32_t> types, widths, offsets;
    for ( auto& a : t.attrs ) {
        int off = s.getOffsetInTuple ( a.name );
        switch ( a.type.tag ) {
            case SqlType::BOOL: types.push_back ( RQ_I8 ); widths.push_back ( 1 ); break;
            case SqlType::INT: case SqlType::DATE: types.push_back ( RQ_I32 ); widths.push_back ( 4 ); break;
            case SqlType::BIGINT: case SqlType::DECIMAL: types.push_back ( RQ_I64 ); widths.push_back ( 8 ); break;
            case SqlType::CHAR:
                if ( a.type.charSpec().num == 1 ) { types.push_back ( RQ_I8 ); widths.push_back ( 1 ); }
                else { types.push_back ( RQ_STR ); widths.push_back ( (int) a.type.charSpec().num + 1 ); }
                break;
            case SqlType::VARCHAR: types.push_back ( RQ_STR ); widths.push_back ( (int) a.type.varcharSpec().num + 1 ); break;
            default: throw ResqlError ( "GPU engine: column type of " + a.name + " not supported" );
        }
        offsets.push_back ( off );
    }
    std::vector<const uint8_t*> blocks;
    std::vector<size_t> sizes;
    /* this rank's row range [lo, hi) of the relation (whole relation unless sharded); blocks hold
     * whole tuples, so a range is a list of (pointer, bytes) pieces of the blocks */
    const size_t tup = s._tupSize;
    const size_t total = t.rel->tupleNum();
    size_t lo = 0, hi = total;
    if ( shard ) { lo = total * (size_t) group().rank / (size_t) group().world; hi = total * (size_t) ( group().rank + 1 ) / (size_t) group().world; }
    size_t first = 0;
    for ( auto& b : t.rel->_dataBlocks ) {
        const size_t n = b->_contentSize / tup;
        const size_t a = std::max ( lo, first ), z = std::min ( hi, first + n );
        if ( a < z ) { blocks.push_back ( b->begin() + ( a - first ) * tup ); sizes.push_back ( ( z - a ) * tup ); }
        first += n;
    }
    rq_table* h = nullptr;
    Timer tm;
    /* gpus=N: a table that is complete on every rank crosses PCIe once, on rank 0, and reaches the other
     * GPUs over NVLink (rq_table_broadcast); all ranks mirror the same tables in the same order */
    const bool replicate = !shard && group().world > 1;
    if ( replicate && group().rank != 0 ) {
        check ( rq_table_alloc ( t.name.c_str(), (int) t.attrs.size(), types.data(), widths.data(), (int64_t) total, &h ) );
    } else
    check ( rq_table_upload_rows ( t.name.c_str(), (int) t.attrs.size(), types.data(), widths.data(), offsets.data(),
                                   (int) s._tupSize, (int) blocks.size(), blocks.data(), sizes.data(), &h ) );
    if ( replicate ) check ( rq_table_broadcast ( h, 0 ) );
    lastLoadMs() += tm.get();
    mirrors() [ key ] = h;
    return h;
}

/* rq_result -> Relation in the reference row format (materialize.h:78-220 output layout) */
static std::unique_ptr<Relation> toRelation ( Schema& schema, rq_result* res ) {
    auto rel = std::make_unique<Relation> ( schema );
    if ( res == nullptr ) return rel;
    auto atts = AttributeIterator::getAll ( rel->_schema );
    Relation::AppendIterator it ( rel.get() );
    for ( int64_t r = 0; r < res->n_rows; r++ ) {
        Data* t = it.get();
        for ( int c = 0; c < res->n_cols; c++ ) {
            rq_result_col& col = res->cols[c];
            Data* dst = atts[c].getPtr ( t );
            const uint8_t* src = (const uint8_t*) col.data + (size_t) r * col.width;
            SqlType ty = atts[c].attribute.type;
            if ( ty.tag == SqlType::CHAR && ty.charSpec().num == 1 ) { dst[0] = src[0]; dst[1] = 0; }
            else memcpy ( dst, src, col.width );
        }
    }
    return rel;
}

}  // namespace rqshim


/* ~HashJoinOp frees _ht unconditionally (hashjoin.h:80-82); the GPU path never allocates the CPU
 * hash table, so hand the destructor a minimal one before the plan is deleted. */
static void rqshimPrepareDelete ( RelOperator* op ) {
    for ( auto* c : op->children ) rqshimPrepareDelete ( c );
    if ( auto* hj = dynamic_cast<HashJoinOp*> ( op ) ) {
        if ( hj->_ht == nullptr ) hj->_ht = allocateHashTable ( (size_t) 1, (size_t) 8 );
    }
}


/**
 * Same contract as executeSelectPlan (execute.h:213-247), executed on the GPU.
 */
std::unique_ptr < SelectResult > executeSelectPlanGpu ( RelOperator*  root,
                                                        bool          requestAll,
                                                        Database&     db,
                                                        DBConfig      config = DBConfig() ) {
    std::stringstream plan;
    if ( config.showPlan ) root->print ( plan );

    ExpressionContext exprCtx;
    root->defineExpressionsForPlan ( exprCtx );
    std::map <std::string, SqlType> identifiers = mapIdentifierTypes ( db );
    exprCtx.deriveExpressionTypes ( identifiers );
    exprCtx.unifyExpressions();

    JitExecutionReport report;
    report.config = config.jit;
    std::unique_ptr<Relation> rel;
    rq_result* res = nullptr;
    try {
        Timer lowerTimer;
        rqshim::Lowering low ( db, requestAll );
        low.produce ( root, {} );
        if ( low.pipelines.empty() || low.pipelines.back().sink_kind != RQ_SINK_MATERIALIZE ) {
            throw ResqlError ( "GPU engine: plan does not end in a materializing operator" );
        }
        report.compilationTime = lowerTimer.get();

        const char* dump = getenv ( "RESQL_B200_DUMP_PLAN" );
        if ( dump != nullptr ) { std::ofstream f ( dump ); f << low.toJson(); }

        if ( getenv ( "RESQL_B200_DRY" ) == nullptr ) {
            rqshim::Group& grp = rqshim::group();
            if ( !rqshim::engineUp() ) {
                const char* dev = getenv ( "RESQL_B200_DEVICE" );
                rqshim::check ( rq_init ( dev ? atoi ( dev ) : grp.rank ) );
                rqshim::engineUp() = true;
                /* engine knobs (rq_set_option) for tests and experiments: RESQL_B200_OPTIONS="key=value,key=value" */
                if ( const char* opts = getenv ( "RESQL_B200_OPTIONS" ) ) {
                    std::string o ( opts );
                    size_t pos = 0;
                    while ( pos < o.size() ) {
                        size_t end = o.find ( ',', pos );
                        if ( end == std::string::npos ) end = o.size();
                        const std::string kv = o.substr ( pos, end - pos );
                        const size_t eq = kv.find ( '=' );
                        if ( eq != std::string::npos ) rqshim::check ( rq_set_option ( kv.substr ( 0, eq ).c_str(), atof ( kv.substr ( eq + 1 ).c_str() ) ) );
                        pos = end + 1;
                    }
                }
            }
            if ( grp.world > 1 && !grp.joined ) {
                /* rank 0 creates the NCCL id and publishes it through a file, the others wait for it */
                uint8_t id[128];
                if ( grp.rank == 0 ) {
                    rqshim::check ( rq_dist_unique_id ( id ) );
                    std::ofstream f ( grp.idFile + ".tmp", std::ios::binary );
                    f.write ( (const char*) id, 128 );
                    f.close();
                    rename ( ( grp.idFile + ".tmp" ).c_str(), grp.idFile.c_str() );
                } else {
                    for ( int tries = 0; ; tries++ ) {
                        std::ifstream f ( grp.idFile, std::ios::binary );
                        if ( f.is_open() && f.read ( (char*) id, 128 ) && f.gcount() == 128 ) break;
                        if ( tries > 60000 ) throw ResqlError ( "GPU engine: rank 0 never published the NCCL id" );
                        usleep ( 1000 );
                    }
                }
                rqshim::check ( rq_dist_init ( grp.rank, grp.world, id ) );
                grp.joined = true;
            }
            /* the fact table of a sharded execution: the table scanned by the pipeline the library
             * merges after (the last aggregation, else the final relation) */
            int shardTable = -1;
            if ( grp.world > 1 ) {
                int mergePipe = (int) low.pipelines.size() - 1;
                for ( size_t i = 0; i < low.pipelines.size(); i++ ) if ( low.pipelines[i].sink_kind == RQ_SINK_AGG ) mergePipe = (int) i;
                if ( low.pipelines[mergePipe].source_kind == RQ_SRC_TABLE ) shardTable = low.pipelines[mergePipe].source_id;
                /* a table that is also scanned by another pipeline must stay complete there */
                for ( size_t i = 0; i < low.pipelines.size(); i++ )
                    if ( (int) i != mergePipe && low.pipelines[i].source_kind == RQ_SRC_TABLE && low.pipelines[i].source_id == shardTable ) shardTable = -1;
            }
            std::vector<rq_table*> handles;
            for ( size_t t = 0; t < low.tables.size(); t++ ) handles.push_back ( rqshim::mirrorTable ( low.tables[t], (int) t == shardTable ) );

            std::vector<rq_pipeline> pls ( low.pipelines.size() );
            for ( size_t i = 0; i < pls.size(); i++ ) {
                rqshim::PipelineDesc& pd = low.pipelines[i];
                rq_pipeline& p = pls[i];
                p.source_kind = pd.source_kind; p.source_id = pd.source_id; p.source_id2 = pd.source_id2; p.reserved = 0;
                p.n_nodes = (int) pd.nodes.size(); p.nodes = pd.nodes.data();
                p.n_args = (int) pd.args.size(); p.args = pd.args.data();
                p.sink_kind = pd.sink_kind;
                p.n_keys = (int) pd.keys.size(); p.keys = pd.keys.data();
                p.n_vals = (int) pd.vals.size(); p.vals = pd.vals.data();
                p.size_hint = pd.size_hint;
            }
            rq_plan rp;
            rp.n_tables = (int) handles.size(); rp.tables = handles.data();
            rp.n_pipelines = (int) pls.size(); rp.pipelines = pls.data();
            rp.n_order = (int) low.order.size(); rp.order = low.order.data();
            rp.limit = low.limit;
            rp.strpool = low.strpool.data(); rp.strpool_bytes = (int64_t) low.strpool.size();
            rp.flags = shardTable >= 0 ? RQ_PLAN_SHARDED : 0;
            rq_timings tm;
            rqshim::check ( rq_plan_execute ( &rp, &res, &tm ) );
            report.compilationTime += tm.lower_ms;
            report.executionTime = tm.kernel_ms;
            if ( config.jit.printPerformance ) {
                std::cout << "gpu: lower " << report.compilationTime << " ms, kernels " << tm.kernel_ms
                          << " ms (scan " << tm.scan_kernel_ms << " ms), d2h " << tm.d2h_ms << " ms, launches "
                          << tm.kernel_launches << ", table load " << rqshim::lastLoadMs() << " ms" << std::endl;
                rqshim::lastLoadMs() = 0;
            }
        }
        rel = rqshim::toRelation ( low.resultSchema, res );
        rq_result_free ( res );
    }
    catch ( ResqlError& err ) {
        std::cerr << err.message();
        rq_result_free ( res );
        rqshimPrepareDelete ( root );
        root->deletePlan();
        throw;
    }
    rqshimPrepareDelete ( root );
    root->deletePlan();
    if ( config.writeResultsToFile ) writeRelationToFile ( *rel, "qres.tbl" );
    return std::make_unique < SelectResult > ( report, std::move ( rel ), plan.str() );
}
