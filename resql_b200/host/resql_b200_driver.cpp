/*
 * resql-b200: non-interactive ReSQL front end whose SELECT statements run on the GPU engine.
 *
 * The reference's lexer/parser/planner/Expr/operator classes are used UNCHANGED (compiled from
 * the reference checkout, see build_host.sh); only the execution span of executeSelectPlan is
 * replaced by executeSelectPlanGpu (gpu_executor.h). Statement dispatch mirrors executeStatement
 * (execute.h:509-543). `engine=cpu` switches back to the reference's own JIT for A/B runs.
 *
 * Usage: resql-b200 [--quiet] STATEMENT...   (same driver statements as the oracle driver:
 *        "out <file>", "repeat <n>", "binload <table> <file>", plus "engine=gpu|cpu" and "gpus=N")
 *
 * gpus=N (must be the first statement; the counterpart of the reference's `threads=N`): the
 * program re-executes itself N-1 times, one process per GPU. Every process runs the same
 * statements on its own copy of the row store; selects run with the fact table row-range sharded
 * (gpu_executor.h) and are merged over NCCL, so every rank computes the full result. Rank 0 prints
 * and writes the output files, the other ranks are silent.
 */
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <fstream>
#include <iostream>
#include <unistd.h>
#include <sys/wait.h>

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

/* gpus=N: become rank 0 and start ranks 1..N-1 as copies of this process */
static std::vector<pid_t> startRanks ( int world, int argc, char** argv ) {
    std::vector<pid_t> kids;
    char idFile[] = "/tmp/resql_b200_nccl_id_XXXXXX";
    int fd = mkstemp ( idFile );
    if ( fd >= 0 ) { close ( fd ); unlink ( idFile ); }
    setenv ( "RESQL_B200_WORLD", std::to_string ( world ).c_str(), 1 );
    setenv ( "RESQL_B200_IDFILE", idFile, 1 );
    for ( int r = 1; r < world; r++ ) {
        pid_t pid = fork();             /* before any CUDA call of this process */
        if ( pid == 0 ) {
            setenv ( "RESQL_B200_RANK", std::to_string ( r ).c_str(), 1 );
            execv ( "/proc/self/exe", argv );
            _exit ( 127 );
        }
        kids.push_back ( pid );
    }
    setenv ( "RESQL_B200_RANK", "0", 1 );
    (void) argc;
    return kids;
}

int main(int argc, char** argv) {
    Database db;
    DBConfig config;
    bool quiet = false, gpu = true;
    std::vector<pid_t> kids;
    if ( getenv ( "RESQL_B200_RANK" ) != nullptr ) {        /* one of the ranks a gpus=N parent started */
        rqshim::group().rank = atoi ( getenv ( "RESQL_B200_RANK" ) );
        rqshim::group().world = atoi ( getenv ( "RESQL_B200_WORLD" ) );
        rqshim::group().idFile = getenv ( "RESQL_B200_IDFILE" );
    }
    const bool silent = rqshim::group().rank != 0;
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
            if (startsWith(st, "gpus=")) {
                const int n = std::stoi(st.substr(5));
                if (getenv("RESQL_B200_RANK") == nullptr && n > 1) {
                    if (rqshim::engineUp()) { std::cout << "Query error: gpus=N must come before the first select" << std::endl; continue; }
                    kids = startRanks(n, argc, argv);
                    rqshim::group().rank = 0;
                    rqshim::group().world = n;
                    rqshim::group().idFile = getenv("RESQL_B200_IDFILE");
                }
                continue;
            }
            if (st == "engine=gpu") { gpu = true; continue; }
            if (st == "engine=cpu") { gpu = false; continue; }
            if (startsWith(st, "out ")) { outFile = st.substr(4); outCount = 0; continue; }
            if (startsWith(st, "repeat ")) { repeat = std::stoi(st.substr(7)); continue; }
            if (startsWith(st, "binload ")) {
                std::string rest = st.substr(8);
                auto sp = rest.find(' ');
                try {
                    size_t n = binload(db, rest.substr(0, sp), rest.substr(sp + 1));
                    if (!silent) std::cout << "Inserted " << n << " tuples" << std::endl;
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
                    if (silent) continue;
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
                if (!silent && (!quiet || res.error)) printQueryResult(res);
            }
        }
    }
    rq_shutdown();
    int rc = 0;
    for (pid_t k : kids) {
        int status = 0;
        waitpid(k, &status, 0);
        if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) rc = 1;
    }
    if (!kids.empty() && getenv("RESQL_B200_IDFILE")) unlink(getenv("RESQL_B200_IDFILE"));
    return rc;
}
