/*
 * Non-interactive driver around the UNMODIFIED reference engine (Henning1/resql).
 *
 * TEST INFRASTRUCTURE ONLY. This TU is compiled by oracle/ref_build/build_ref.sh against the
 * reference headers where they lie under /root/reference (patched copy in a /tmp build dir, see
 * the script) and produces oracle/_ref/resql-oracle. It replaces src/resql.cpp (which needs
 * readline headers that this image lacks) and drives the reference exclusively through its own
 * public statement API: expandExecStatements / executeStatement / printQueryResult
 * (execute.h:477, :509, :173) and serializeRelation (dbdata.h:688).
 *
 * Usage:  resql-oracle [--quiet] STATEMENT...
 * Every argument is one reference statement ("exec file.sql", "threads=4", "select ...") or one
 * of these driver-only statements:
 *   out <file>          write the full result of every following select to <file> in the
 *                       reference's `tofile` format (serializeRelation), one file per select:
 *                       <file>, then <file>.1, <file>.2, ...
 *   repeat <n>          run every following select n times (timing lines for each run)
 *   binload <table> <file>   append packed tuples (exactly Schema::_tupSize bytes each, the
 *                       reference's row format) from a binary file through
 *                       Relation::AppendIterator (dbdata.h:217-301) - avoids the leaky text loader
 * For every select a machine-readable line is printed:
 *   #select rows=<n> compile_ms=<c> execute_ms=<e>
 */
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <fstream>
#include <iostream>

#include "operators/JitOperators.h"
#include "execute.h"

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
            Data* dst = it.get();
            memcpy(dst, buf.data() + o, tup);
            n++;
        }
    }
    return n;
}

int main(int argc, char** argv) {
    Database db;
    DBConfig config;
    bool quiet = false;
    std::string outFile;
    int outCount = 0;
    int repeat = 1;
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
                QueryResult res = executeStatement(st, db, config);
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
    return 0;
}
