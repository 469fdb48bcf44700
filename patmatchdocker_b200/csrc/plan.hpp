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

enum PosOp : unsigned char { OP_NONE = 0, OP_OPT = 1, OP_STAR = 2, OP_PLUS = 3 };   // position followed by '?', '*', '+'

struct Pattern {
    std::vector<ByteSet> pos;          // one class per pattern position
    std::vector<unsigned char> op;     // PosOp per position (all OP_NONE for SIMPLE / ESIMPLE patterns)
    bool start_line = false, end_line = false;
    bool had_ops = false;              // the pattern was written with ? * + (possibly all rewritten away by the parser rules)
    int m() const { return (int)pos.size(); }
    bool extended() const { for (unsigned char o : op) if (o) return true; return false; }
    bool optional(int j) const { return op[j] == OP_OPT || op[j] == OP_STAR; }
    bool repeats(int j) const { return op[j] == OP_STAR || op[j] == OP_PLUS; }
};

struct Options { int k = 0; bool ins = true, del = true, subs = true; };

enum PlanType { SIMPLE = 0, SPLIT = 1, BWD = 2, FWD = 3,
                EXT_BEG = 4,          // EXTENDED pattern, verification anchored at the START of the scanned sub-pattern
                EXT_END = 5 };        // ... at its END (extendedPreproc @413260: types 2 and 3 of that engine)

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
    // EXTENDED plans (k = 0): what extendedFindBest @411fe0 chose, and the run of plain positions around the
    // anchor that the exact scan looks for (every match contains it at a fixed offset from the anchor)
    int ext_beg = 0, ext_end = 0, ext_wlen = 0, anchor = 0;
    int ext_repeats = 0;               // some position carries '*' or '+': hits have no length bound
    int win_lo = 0, win_hi = 0;
    // closure masks of the two verification walks (extendedLoadVerif @412c60), elements numbered away from the anchor
    uint64_t IL = 0, FL = 0, AL = 0, initL = 0, IR = 0, FR = 0, AR = 0, initR = 0;
    int ext_lead_opt = 0;              // EXT_END plan whose pattern begins with an optional position (scan quirk, see plan.cpp)
    uint64_t IS = 0, FS = 0, AS = 0;   // closure masks of the forward scan over P[0, anchor)
};

// error codes follow include/patmatch_b200.h
int parse_kopt(const char *kopt, Options &o, std::string &err);
int parse_pattern(const char *pattern, bool icase, Pattern &P, std::string &err);
int make_plan(const Pattern &P, const Options &o, Plan &plan, std::string &err);
// process-wide: piece choice of the deployed binary (default glibc allocator) instead of the defined zero-scratch behaviour
void set_compat_deployed(int on);
int compat_deployed();

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
