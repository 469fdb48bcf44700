// plan.hpp -- host-side pattern compiler and search planner of the product path.
//
// Mirrors what the reference engine does before it scans (nrgrep_coords:
// main @400e00 option parsing, parse @41ae90, esimplePreproc @415540,
// simpleFindBest @416a10, esimpleLoadFast @415370, simpleLoadVerif @4173f0),
// producing the tables the CUDA kernels consume.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <array>

namespace pm {

struct ByteSet {                       // 256-bit byte class of one pattern position
    uint64_t w[4] = {0, 0, 0, 0};
    bool has(unsigned c) const { return (w[c >> 6] >> (c & 63)) & 1; }
    void add(unsigned c) { w[c >> 6] |= 1ULL << (c & 63); }
    void invert() { for (auto &x : w) x = ~x; }
    void fill() { for (auto &x : w) x = ~0ULL; }
};

struct Pattern {
    std::vector<ByteSet> pos;          // one class per pattern position
    bool start_line = false, end_line = false;
    int m() const { return (int)pos.size(); }
};

struct Options { int k = 0; bool ins = true, del = true, subs = true; };

enum PlanType { SIMPLE = 0, SPLIT = 1, BWD = 2, FWD = 3 };

struct Plan {
    PlanType type = SIMPLE;
    int m = 0, k = 0;
    bool ins = true, del = true, subs = true;
    int L = 0;                         // piece length (SPLIT) / scanned sub-pattern length (BWD, FWD) / m (SIMPLE)
    int npieces = 1;
    int V[16] = {0};                   // split points
    uint64_t trig[16] = {0};           // SPLIT: mask over piece-top bits that make piece i a candidate
    double split_cost = 0, fb_cost = 0;
    int fb_flag = 0, fb_beg = 0, fb_end = 0;
};

// error codes follow include/patmatch_b200.h
int parse_kopt(const char *kopt, Options &o, std::string &err);
int parse_pattern(const char *pattern, bool icase, Pattern &P, std::string &err);
int make_plan(const Pattern &P, const Options &o, Plan &plan, std::string &err);

// Device tables ---------------------------------------------------------------
struct FilterTables {                  // byte Shift-And over the superimposed pieces
    uint64_t B[256];                   // bit i*L+j set iff piece i position j accepts the byte
    uint64_t init;                     // bits i*L
    uint64_t fin;                      // bits i*L+L-1
    int bits;                          // npieces*L
};
struct VerifyTables {                  // per piece: anchored NFA masks, simpleLoadVerif @4173f0
    std::vector<uint64_t> TL, TR;      // [piece][256]: left part read leftwards / right part read rightwards
};
void build_filter(const Pattern &P, const Plan &plan, FilterTables &ft);
void build_verify(const Pattern &P, const Plan &plan, VerifyTables &vt);

}  // namespace pm
