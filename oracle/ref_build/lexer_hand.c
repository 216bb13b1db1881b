/*
 * Hand-written SQL tokenizer standing in for the flex output of the reference's
 * src/parser/lexer.y (flex is not installed in this image).
 *
 * TEST INFRASTRUCTURE ONLY: this file is compiled into the reference oracle binary
 * (oracle/_ref/resql-oracle); it is never linked into the product library.
 *
 * Behaviour follows lexer.y rule by rule: longest match, earlier rule wins ties
 * (lexer.y:19-241). It exports exactly the symbols parseSql.h:16-23 expects:
 * yytext, yyin, yylex, yy_scan_string, yy_delete_buffer.
 * Token codes come from the lemon-generated parser.h; ERROR from parser/common.h:1.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "parser.h"
#define ERROR 99999

char* yytext = NULL;
FILE* yyin = NULL;

typedef struct yy_buffer_state { char* buf; size_t len; size_t pos; } *YY_BUFFER_STATE;
static YY_BUFFER_STATE cur = NULL;
static char* tokbuf = NULL;
static size_t tokcap = 0;

YY_BUFFER_STATE yy_scan_string(const char* s) {
    YY_BUFFER_STATE b = (YY_BUFFER_STATE)malloc(sizeof(*b));
    b->len = strlen(s);
    b->buf = (char*)malloc(b->len + 1);
    memcpy(b->buf, s, b->len + 1);
    b->pos = 0;
    cur = b;
    return b;
}

void yy_delete_buffer(YY_BUFFER_STATE b) {
    if (!b) return;
    if (cur == b) cur = NULL;
    free(b->buf);
    free(b);
}

static int is_digit(char c) { return c >= '0' && c <= '9'; }
static int is_idstart(char c) { return c >= 'a' && c <= 'z'; }
static int is_idchar(char c) { return is_idstart(c) || is_digit(c) || c == '_'; }

/* literal rules in lexer.y file order (order breaks equal-length ties, and all
 * of them precede the IDENTIFIER rule) */
static const struct { const char* text; int tok; } kLiterals[] = {
    {"select", SELECT_TK}, {"from", FROM}, {"where", WHERE}, {"group by", GROUPBY},
    {"order by", ORDERBY}, {"limit", LIMIT_TK}, {"asc", ASC_TK}, {"desc", DESC_TK},
    {"create table", CREATE_TABLE_TK}, {"bulk insert", BULK_INSERT_TK},
    {"fieldterminator", FIELDTERMINATOR_TK}, {"firstrow", FIRSTROW_TK}, {"with", WITH_TK},
    {"sum", SUM_TK}, {"count", COUNT_TK}, {"avg", AVG_TK}, {"min", MIN_TK}, {"max", MAX_TK},
    {"between", BETWEEN_TK}, {"(", LPAREN}, {")", RPAREN}, {"+", PLUS_TK}, {"-", MINUS_TK},
    {"*", MUL_TK}, {"/", DIV_TK}, {">=", GE_TK}, {">", GT_TK}, {"<=", LE_TK}, {"<", LT_TK},
    {"=", EQ_TK}, {"<>", NEQ_TK}, {",", COMMA}, {"::", TYPECAST_TK}, {"and", AND_TK},
    {"in", IN_TK}, {"like", LIKE_TK}, {"or", OR_TK}, {"as", AS_TK}, {"bigint", BIGINT_TK},
    {"int", INT_TK}, {"date", DATE_TK}, {"decimal", DECIMAL_TK}, {"char", CHAR_TK},
    {"varchar", VARCHAR_TK}, {"case", CASE_TK}, {"when", WHEN_TK}, {"then", THEN_TK},
    {"else", ELSE_TK}, {"end", END_TK},
};

/* number rules: returns match length and token, 0 if none */
static size_t match_number(const char* p, size_t n, int* tok) {
    size_t i = 0, intd = 0, fracd = 0;
    int dot = 0;
    while (i < n && is_digit(p[i])) { i++; intd++; }
    if (i < n && p[i] == '.') {
        size_t j = i + 1, f = 0;
        while (j < n && is_digit(p[j])) { j++; f++; }
        if (intd > 0 || f > 0) { dot = 1; fracd = f; i = j; }
    }
    if (intd == 0 && !(dot && fracd > 0)) return 0;
    size_t mant = i;
    /* optional exponent turns any mantissa form into FLOAT */
    if (mant < n && p[mant] == 'e') {
        size_t j = mant + 1;
        if (j < n && (p[j] == '+' || p[j] == '-')) j++;
        size_t e = 0;
        while (j < n && is_digit(p[j])) { j++; e++; }
        if (e > 0) { *tok = FLOAT_CONSTANT; return j; }
    }
    *tok = dot ? DECIMAL_CONSTANT : INTEGER_CONSTANT;
    return mant;
}

static size_t match_string(const char* p, size_t n) {
    if (n == 0 || (p[0] != '"' && p[0] != '\'')) return 0;
    char q = p[0];
    size_t i = 1;
    while (i < n) {
        if (p[i] == '\\') { if (i + 1 >= n) return 0; i += 2; continue; }
        if (p[i] == q) return i + 1;
        i++;
    }
    return 0;
}

int yylex(void) {
    if (!cur) return 0;
    for (;;) {
        if (cur->pos >= cur->len) return 0;
        const char* p = cur->buf + cur->pos;
        size_t n = cur->len - cur->pos;
        /* whitespace */
        if (*p == ' ' || *p == '\t' || *p == '\n') { cur->pos++; continue; }
        /* one-line comment: needs the terminating newline to match (lexer.y:233) */
        if (n >= 2 && p[0] == '-' && p[1] == '-') {
            const char* nl = memchr(p, '\n', n);
            if (nl) { cur->pos += (size_t)(nl - p) + 1; continue; }
        }
        size_t best = 0;
        int tok = ERROR;
        int t;
        size_t l = match_number(p, n, &t);
        if (l > best) { best = l; tok = t; }
        l = match_string(p, n);
        if (l > best) { best = l; tok = STRING_CONSTANT; }
        for (size_t k = 0; k < sizeof(kLiterals) / sizeof(kLiterals[0]); k++) {
            size_t ll = strlen(kLiterals[k].text);
            if (ll <= n && ll > best && memcmp(p, kLiterals[k].text, ll) == 0) {
                best = ll; tok = kLiterals[k].tok;
            }
        }
        if (is_idstart(*p)) {
            size_t i = 1;
            while (i < n && is_idchar(p[i])) i++;
            if (i > best) { best = i; tok = IDENTIFIER; }
        }
        if (best == 0) {
            best = 1;
            tok = ERROR;
        }
        if (best + 1 > tokcap) { tokcap = best + 64; tokbuf = (char*)realloc(tokbuf, tokcap); }
        memcpy(tokbuf, p, best);
        tokbuf[best] = 0;
        yytext = tokbuf;
        cur->pos += best;
        if (tok == ERROR) printf("Unrecognized character: %s\n", yytext);
        return tok;
    }
}
