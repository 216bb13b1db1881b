/*
 * resql-b200: non-interactive ReSQL front end whose SELECT statements run on the GPU engine.
 *
 * The reference's lexer/parser/planner/Expr/operator classes are used UNCHANGED (compiled from
 * the reference checkout, see build_host.sh); only the execution span of executeSelectPlan is
 * replaced by executeSelectPlanGpu (gpu_executor.h). Statement dispatch mirrors executeStatement
 * (execute.h:509-543). `engine=cpu` switches back to the reference's own JIT for A/B runs.
 *
 * Usage: resql-b200 [--quiet] STATEMENT...   (same driver statements as the oracle driver:
 *        "out <file>", "repeat <n>", "binload <table> <file>", plus "engine=gpu|cpu")
 */
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <fstream>
#include <iostream>

#include "operators/JitOperators.h"
#include "execute.h"
#include "gpu_executor.h"

size_t DataBlock::Size = 2 << 20;

static bool startsWith(const std::string& s, const char* p) { return s.rfind(p, 0) == 0; }

static size_t binload(Database& db, const std::string& table, const std::string& file) {
    if (db.relations.count(table) == 0) throw ResqlError("binload: no table " + table);
    Relation& rel = db.relations[table];
    size_t tup = rel._schema._tupSize;
    std::ifstream f(file, std::ios::binary);
    if (!f.is_open()) throw ResqlError("binload: cannot open " + file);
    auto it = Relation::AppendIterator(&rel);
    std::vector<char> buf(tup * 4096);
    size_t n = 0;
    while (f) {
        f.read(buf.data(), buf.size());
        size_t got = (size_t)f.gcount();
        for (size_t o = 0; o + tup <= got; o += tup) {
            memcpy(it.get(), buf.data() + o, tup);
            n++;
        }
    }
    return n;
}

static QueryResult executeStatementGpu(std::string statement, Database& db, DBConfig& config, bool gpu) {
    if (!gpu) return executeStatement(statement, db, config);
    try {
        ControlResult res = processControl(statement, db, config);
        if (res.actionDone) return res;
        Query query = parseSql(statement);
        if (query.parseError) throw ResqlError("Syntax error.");
        if (query.tag == Query::SELECT) {
            buildQuery(query, db);
            if (query.plan == nullptr) throw ResqlError("Could not generate query plan.");
            return executeSelectPlanGpu(query.plan, query.requestAll, db, config);
        } else if (query.tag == Query::CREATE_TABLE) {
            return executeCreateTable(query, db);
        } else if (query.tag == Query::BULK_INSERT) {
            return executeBulkInsert(query, db);
        }
    } catch (std::runtime_error& e) {
        return QueryResult(e);
    } catch (ResqlError& e) {
        return QueryResult(e);
    }
    return ResqlError("unwanted fallthrough");
}

int main(int argc, char** argv) {
    Database db;
    DBConfig config;
    bool quiet = false, gpu = true;
    std::string outFile;
    int outCount = 0, repeat = 1;
    for (int i = 1; i < argc; i++) {
        std::string arg = argv[i];
        if (arg == "--quiet") { quiet = true; continue; }
        std::vector<std::string> statements;
        try {
            statements = expandExecStatements(arg);
        } catch (ResqlError& e) {
            std::cout << "Query error: " << e.message() << std::endl;
            continue;
        }
        for (auto& s : statements) {
            std::string st = s;
            rtrim(st); ltrim(st);
            if (st == "engine=gpu") { gpu = true; continue; }
            if (st == "engine=cpu") { gpu = false; continue; }
            if (startsWith(st, "out ")) { outFile = st.substr(4); outCount = 0; continue; }
            if (startsWith(st, "repeat ")) { repeat = std::stoi(st.substr(7)); continue; }
            if (startsWith(st, "binload ")) {
                std::string rest = st.substr(8);
                auto sp = rest.find(' ');
                try {
                    size_t n = binload(db, rest.substr(0, sp), rest.substr(sp + 1));
                    std::cout << "Inserted " << n << " tuples" << std::endl;
                } catch (ResqlError& e) {
                    std::cout << "Query error: " << e.message() << std::endl;
                }
                continue;
            }
            int reps = 1;
            for (int r = 0; r < reps; r++) {
                QueryResult res = executeStatementGpu(st, db, config, gpu);
                if (!res.error && res.tag == Query::SELECT) {
                    reps = repeat;
                    SelectResult* sel = res.selectResult();
                    std::cout << "#select rows=" << sel->relation->tupleNum()
                              << " compile_ms=" << sel->jitReport.compilationTime
                              << " execute_ms=" << sel->jitReport.executionTime << std::endl;
                    if (!outFile.empty() && r == 0) {
                        std::string fn = outFile;
                        if (outCount > 0) fn += "." + std::to_string(outCount);
                        outCount++;
                        std::ofstream f(fn);
                        f << "#schema";
                        for (auto& a : sel->relation->_schema._attribs)
                            f << " " << a.name << ":" << serializeType(a.type);
                        f << "\n";
                        serializeRelation(*sel->relation, f);
                    }
                }
                if (!quiet || res.error) printQueryResult(res);
            }
        }
    }
    rq_shutdown();
    return 0;
}
