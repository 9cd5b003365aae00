// eqtlbma_bf -- drop-in host front-end of the B200-native hot path.
//
// Same command line, same input files and same gzipped outputs as the reference's eqtlbma_bf
// (timflutre/eqtlbma v1.3.3, src/eqtlbma_bf.cpp): this program parses the inputs into the flat
// all-sample-space layouts of include/eqtlbma_b200.h with the reference loader's ordering rules
// (SURVEY.md App. B #1), hands the hot path -- testForAssociations / makePermutations,
// eqtlbma_bf.cpp:1546-1574 -- to libeqtlbma_b200.so through its C ABI, and serialises the results
// with the reference's text conventions (writeRes*, eqtlbma_bf.cpp:919-1447).  No statistics are
// computed on the host.  Out of scope here (reported as errors): --lik poisson /
// quasipoisson, tabix-indexed --scoord (an index is ignored: every SNP of the BED
// file is loaded, which gives the same cis sets).
#include <getopt.h>
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <ctime>
#include <deque>
#include <iostream>
#include <limits>
#include <map>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <dirent.h>
#include <sys/time.h>
#include <sys/wait.h>
#include <thread>
#include <unistd.h>
#include <vector>

#include "../../include/eqtlbma_b200.h"

using namespace std;

#ifndef EQB_VERSION
#define EQB_VERSION "1.3.3-b200"
#endif

namespace {

const double kNaN = numeric_limits<double>::quiet_NaN();

// The CUDA context is created on a thread while the input files are parsed; every exit path waits for that thread first
// (the CUDA runtime must not be torn down by exit() while another thread is still inside its initialisation).
std::thread *g_warm = NULL;
[[noreturn]] void eqb_exit(int code)
{
  if (g_warm && g_warm->joinable()) g_warm->join();
  exit(code);
}

// ------------------------------------------------------------------ small text / gz helpers
// utils::split with strtok semantics: consecutive delimiters collapse (utils_io.cpp:48-62)
void split(const string &s, const char *delim, vector<string> &tokens)
{
  tokens.clear();
  size_t i = 0, n = s.size();
  while (i < n) {
    while (i < n && strchr(delim, s[i])) ++i;
    if (i >= n) break;
    size_t j = i;
    while (j < n && !strchr(delim, s[j])) ++j;
    tokens.push_back(s.substr(i, j - i));
    i = j;
  }
}

// The two big matrices (custom-format genotypes, expression levels) are tokenised in place -- no std::string per cell --
// and their cells go through a decimal fast path: a literal [-]ddd[.ddd] with <= 15 significant digits is m / 10^d with m
// and 10^d exactly representable, and the IEEE quotient of two exact doubles is the correctly rounded value of the
// decimal, i.e. the very double strtod / atof returns (Clinger's fast path).  Everything else (exponents, inf, nan, long
// mantissas) goes through strtod.  data_loader.cpp:437-527, 878-1010 parse the same cells with utils::split + atof.
struct Span {
  const char *p;
  size_t n;
  bool eq(const char *s) const { return strlen(s) == n && memcmp(p, s, n) == 0; }
  string str() const { return string(p, n); }
};

void split_spans(const string &s, vector<Span> &tok) // delimiters: space and tab (strtok semantics)
{
  tok.clear();
  const char *b = s.data(), *e = b + s.size();
  while (b < e) {
    while (b < e && (*b == ' ' || *b == '\t')) ++b;
    if (b >= e) break;
    const char *t = b;
    while (b < e && *b != ' ' && *b != '\t') ++b;
    Span sp;
    sp.p = t;
    sp.n = (size_t)(b - t);
    tok.push_back(sp);
  }
}

bool is_na(const Span &t)
{
  if (t.n < 2 || t.n > 3 || (t.p[0] != 'N' && t.p[0] != 'n')) return false; // (a number never starts with n)
  return t.eq("NA") || t.eq("na") || t.eq("NaN") || t.eq("nan");
}

double fast_atof(const Span &t)
{
  static const double p10[] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                               1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  const char *b = t.p, *e = t.p + t.n;
  bool neg = false;
  if (b < e && (*b == '-' || *b == '+')) neg = (*b++ == '-');
  unsigned long long m = 0;
  int nd = 0, dec = 0;
  bool any = false, ok = true;
  while (b < e && *b >= '0' && *b <= '9') {
    if (m || *b != '0') ++nd;
    m = m * 10 + (unsigned long long)(*b++ - '0');
    any = true;
    if (nd > 15) { ok = false; break; }
  }
  if (ok && b < e && *b == '.') {
    ++b;
    while (b < e && *b >= '0' && *b <= '9') {
      if (m || *b != '0') ++nd;
      m = m * 10 + (unsigned long long)(*b++ - '0');
      ++dec;
      any = true;
      if (nd > 15 || dec > 22) { ok = false; break; }
    }
  }
  if (ok && any && b == e) {
    const double v = (double)m / p10[dec];
    return neg ? -v : v;
  }
  char buf[64];
  if (t.n < sizeof(buf)) {
    memcpy(buf, t.p, t.n);
    buf[t.n] = 0;
    return atof(buf);
  }
  return atof(t.str().c_str());
}

// Line reader over a gzip (or plain) file.  zlib inflates on a background thread into 1 MiB blocks (a bounded queue), the
// caller's thread cuts lines out of them: decompression -- half the loading time of a dosage matrix -- overlaps the parsing.
struct GzReader {
  gzFile f;
  string path;
  std::thread th;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<string> q;
  bool done = false, stop = false, started = false;
  string cur;
  size_t pos = 0;
  explicit GzReader(const string &p) : path(p)
  {
    f = gzopen(p.c_str(), "rb");
    if (f == NULL) {
      cerr << "ERROR: can't open file " << p << " with mode rb" << endl;
      eqb_exit(EXIT_FAILURE);
    }
    gzbuffer(f, 1 << 20);
  }
  void producer()
  {
    const size_t BLOCK = 1 << 20, MAXQ = 6;
    while (true) {
      string b(BLOCK, '\0');
      const int n = gzread(f, &b[0], (unsigned)BLOCK);
      std::unique_lock<std::mutex> lk(mu);
      if (n <= 0) {
        done = true;
        cv.notify_all();
        return;
      }
      b.resize((size_t)n);
      cv.wait(lk, [&] { return q.size() < MAXQ || stop; });
      if (stop) return;
      q.push_back(std::move(b));
      cv.notify_all();
    }
  }
  bool next_block()
  {
    if (!started) {
      started = true;
      th = std::thread(&GzReader::producer, this);
    }
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return !q.empty() || done; });
    if (q.empty()) return false;
    cur = std::move(q.front());
    q.pop_front();
    pos = 0;
    cv.notify_all();
    return true;
  }
  // true for every line ended by a newline (possibly empty) and for a non-empty last line without one
  bool getline(string &line)
  {
    line.clear();
    bool got = false;
    while (true) {
      if (pos >= cur.size() && !next_block()) return got && !line.empty();
      const char *b = cur.data() + pos;
      const char *nl = (const char *)memchr(b, '\n', cur.size() - pos);
      if (nl) {
        line.append(b, (size_t)(nl - b));
        pos += (size_t)(nl - b) + 1;
        return true;
      }
      line.append(b, cur.size() - pos);
      pos = cur.size();
      got = true;
    }
  }
  ~GzReader()
  {
    if (started) {
      {
        std::unique_lock<std::mutex> lk(mu);
        stop = true;
        cv.notify_all();
      }
      th.join();
    }
    gzclose(f);
  }
};

bool file_exists(const string &p)
{
  FILE *f = fopen(p.c_str(), "rb");
  if (f) fclose(f);
  return f != NULL;
}

bool is_na(const string &t) { return t == "NA" || t == "na" || t == "NaN" || t == "nan"; }

// ostream << double with precision(6) + scientific (eqtlbma_bf.cpp:953-954)
void put_sci(string &out, double v)
{
  char b[64];
  snprintf(b, sizeof(b), "%.6e", v);
  out += b;
}

// ostream << double in the default format (precision 6)
void put_def(string &out, double v)
{
  char b[64];
  snprintf(b, sizeof(b), "%g", v);
  out += b;
}

void gz_write(const string &path, const char *mode, const string &txt)
{
  gzFile f = gzopen(path.c_str(), mode);
  if (f == NULL) {
    cerr << "ERROR: can't open file " << path << " with mode " << mode << endl;
    eqb_exit(EXIT_FAILURE);
  }
  gzbuffer(f, 1 << 20);
  size_t off = 0;
  while (off < txt.size()) {
    const unsigned chunk = (unsigned)min<size_t>(txt.size() - off, 1u << 30);
    if (gzwrite(f, txt.data() + off, chunk) <= 0) {
      cerr << "ERROR: can't write to file " << path << endl;
      eqb_exit(EXIT_FAILURE);
    }
    off += chunk;
  }
  gzclose(f);
}

// One gzip member from a text chunk (in memory).  Concatenated members are a valid gzip file -- it is
// also what the reference produces by re-opening its outputs in "ab" mode for every write-group
// (eqtlbma_bf.cpp:1121-1125).
void deflate_member(const string &txt, vector<unsigned char> &out)
{
  out.clear();
  if (txt.empty()) return;
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) {
    cerr << "ERROR: deflateInit2 failed" << endl;
    eqb_exit(EXIT_FAILURE);
  }
  out.resize(deflateBound(&zs, txt.size()) + 64);
  zs.next_in = (Bytef *)txt.data();
  zs.avail_in = (uInt)txt.size();
  zs.next_out = out.data();
  zs.avail_out = (uInt)out.size();
  if (deflate(&zs, Z_FINISH) != Z_STREAM_END) {
    cerr << "ERROR: deflate failed" << endl;
    eqb_exit(EXIT_FAILURE);
  }
  out.resize(zs.total_out);
  deflateEnd(&zs);
}

// Output encoder: the genes [g0, g1) of a batch are cut into `nthreads` contiguous slices;
// every slice is formatted and deflated by its own thread, the members are appended in gene order.
// fmt(ga, gb, txt) must append the text of genes [ga, gb) to txt.
template <class F>
void parallel_emit(const string &path, int nthreads, int64_t g0, int64_t g1, const vector<int64_t> &weight_prefix, F fmt)
{
  nthreads = max(1, nthreads);
  // slice boundaries balanced on the number of pairs (weight_prefix[g - g0] = pairs before gene g)
  vector<int64_t> cut(nthreads + 1, g1);
  cut[0] = g0;
  const int64_t total = weight_prefix[g1 - g0];
  for (int t = 1; t < nthreads; ++t) {
    const int64_t target = total * t / nthreads;
    int64_t g = cut[t - 1];
    while (g < g1 && weight_prefix[g - g0] < target) ++g;
    cut[t] = g;
  }
  vector<vector<unsigned char> > members(nthreads);
  vector<thread> pool;
  auto work = [&](int t) {
    // chunks of at most ~256 MB of text per member (zlib's 32-bit counters)
    string txt;
    fmt(cut[t], cut[t + 1], txt);
    if (txt.size() < ((size_t)1 << 31))
      deflate_member(txt, members[t]);
    else {
      size_t off = 0;
      vector<unsigned char> part;
      while (off < txt.size()) {
        size_t len = min<size_t>((size_t)1 << 28, txt.size() - off);
        const size_t nl = txt.rfind('\n', off + len - 1);
        if (nl != string::npos && nl >= off) len = nl + 1 - off;
        deflate_member(txt.substr(off, len), part);
        members[t].insert(members[t].end(), part.begin(), part.end());
        off += len;
      }
    }
  };
  for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
  work(0);
  for (size_t t = 0; t < pool.size(); ++t) pool[t].join();
  FILE *f = fopen(path.c_str(), "ab");
  if (f == NULL) {
    cerr << "ERROR: can't open file " << path << " with mode ab" << endl;
    eqb_exit(EXIT_FAILURE);
  }
  for (int t = 0; t < nthreads; ++t)
    if (!members[t].empty() && fwrite(members[t].data(), 1, members[t].size(), f) != members[t].size()) {
      cerr << "ERROR: can't write to file " << path << endl;
      eqb_exit(EXIT_FAILURE);
    }
  fclose(f);
}

// ------------------------------------------------------------------ options (eqtlbma_bf.cpp:1586-1597)
struct Options {
  int verbose = 1, trick = 0, perm_sep = 0, nb_threads = 1, wrtsize = 10;
  size_t radius = 100000, nb_permutations = 0, seed = string::npos, tricut = 10;
  float min_maf = 0.0, fiterr = 0.5;
  bool outss = false, outw = false, qnorm = false, maxbf = false;
  string geno, scoord, exp, gcoord, anchor = "TSS", inss, out, lik = "normal", analys, covar, gridL, gridS,
      bfs = "gen", error = "uvlr", pbf = "none", snp;
  vector<string> sbgrp;
  int device = 0;
  int shard_k = 0, shard_n = 1; // --shard k/N (extension): this process handles the k-th of N gene shards
  int gpus = 1;                 // --gpus N (extension): launcher, one shard process per GPU + merge of the outputs
  bool shard_child = false;     // (internal) process started by the --gpus launcher
  bool no_header = false;       // (internal) shards after the first one write no header lines
};

void help(char **argv)
{
  cout << "`" << argv[0] << "' performs eQTL mapping in multiple subgroups via a Bayesian model" << endl
       << "(B200-native hot path; same options as eqtlbma_bf 1.3.3)." << endl
       << endl
       << "Usage: " << argv[0] << " [OPTIONS] ..." << endl
       << endl
       << "  -h, --help / -V, --version / -v, --verbose" << endl
       << "      --geno --scoord --exp --gcoord --anchor --cis --out --outss --outw --lik --analys" << endl
       << "      --qnorm --maf --covar --gridL --gridS --bfs --error --fiterr --nperm --seed --trick" << endl
       << "      --tricut --permsep --pbf --maxbf --thread --snp --sbgrp --wrtsize" << endl
       << "      --device\tCUDA device ordinal (extension)" << endl
       << "      --shard\tk/N: handle the k-th of N contiguous, cost-balanced gene shards (extension;\n"
       << "\t\tone process per GPU, outputs concatenate in shard order like the batches of\n"
       << "\t\teqtlbma_bf_parallel.bash)" << endl
       << "      --gpus\tN: run the N shards on GPUs 0..N-1 of this machine, one process each, and\n"
       << "\t\tconcatenate their outputs in shard order (extension; replaces the launcher +\n"
       << "\t\t`zcat | sed 1d` merge of eqtlbma_bf_parallel.bash; same files as a single run)" << endl
       << endl
       << "Limits of the device path: at most 2048 samples in the union over the subgroups, 64 subgroups\n"
       << "(20 with --bfs all, 16 with --error mvlr|hybrid; --inss: 10 with --bfs all); --error hybrid: at most 6 covariates,\n"
       << "covariate files in the order of the sorted sample names; not built: --lik other than\n"
       << "normal, tabix-indexed --scoord." << endl;
}

void die_usage(int argc, char **argv, const string &msg)
{
  cerr << "cmd-line:";
  for (int i = 0; i < argc; ++i) cerr << " " << argv[i];
  cerr << endl << endl << "ERROR: " << msg << endl << endl;
  help(argv);
  eqb_exit(EXIT_FAILURE);
}

void parse_cmdline(int argc, char **argv, Options &o)
{
  static struct option long_options[] = {
      {"help", no_argument, 0, 'h'},          {"version", no_argument, 0, 'V'},
      {"verbose", required_argument, 0, 'v'}, {"geno", required_argument, 0, 0},
      {"scoord", required_argument, 0, 0},    {"exp", required_argument, 0, 0},
      {"gcoord", required_argument, 0, 0},    {"anchor", required_argument, 0, 0},
      {"cis", required_argument, 0, 0},       {"inss", required_argument, 0, 0},
      {"out", required_argument, 0, 0},       {"outss", no_argument, 0, 0},
      {"outm", no_argument, 0, 0},            {"outw", no_argument, 0, 0},
      {"lik", required_argument, 0, 0},       {"analys", required_argument, 0, 0},
      {"qnorm", no_argument, 0, 0},           {"maf", required_argument, 0, 0},
      {"covar", required_argument, 0, 0},     {"gridL", required_argument, 0, 0},
      {"gridS", required_argument, 0, 0},     {"bfs", required_argument, 0, 0},
      {"error", required_argument, 0, 0},     {"fiterr", required_argument, 0, 0},
      {"nperm", required_argument, 0, 0},     {"seed", required_argument, 0, 0},
      {"trick", required_argument, 0, 0},     {"tricut", required_argument, 0, 0},
      {"permsep", required_argument, 0, 0},   {"pbf", required_argument, 0, 0},
      {"maxbf", no_argument, 0, 0},           {"thread", required_argument, 0, 0},
      {"snp", required_argument, 0, 0},       {"sbgrp", required_argument, 0, 0},
      {"wrtsize", required_argument, 0, 0},   {"device", required_argument, 0, 0},
      {"shard", required_argument, 0, 0},     {"gpus", required_argument, 0, 0},
      {0, 0, 0, 0}};
  while (true) {
    int idx = 0;
    const int c = getopt_long(argc, argv, "hVv:", long_options, &idx);
    if (c == -1) break;
    if (c == 'h') {
      help(argv);
      eqb_exit(0);
    }
    if (c == 'V') {
      cout << argv[0] << " " << EQB_VERSION << endl;
      eqb_exit(0);
    }
    if (c == 'v') {
      o.verbose = atoi(optarg);
      continue;
    }
    if (c != 0) {
      printf("\n");
      help(argv);
      abort();
    }
    const string n = long_options[idx].name;
    if (n == "geno") o.geno = optarg;
    else if (n == "scoord") o.scoord = optarg;
    else if (n == "exp") o.exp = optarg;
    else if (n == "gcoord") o.gcoord = optarg;
    else if (n == "anchor") o.anchor = optarg;
    else if (n == "cis") o.radius = atol(optarg);
    else if (n == "inss") o.inss = optarg;
    else if (n == "out") o.out = optarg;
    else if (n == "outss") o.outss = true;
    else if (n == "outw") o.outw = true;
    else if (n == "lik") o.lik = optarg;
    else if (n == "analys") o.analys = optarg;
    else if (n == "qnorm") o.qnorm = true;
    else if (n == "maf") o.min_maf = atof(optarg);
    else if (n == "covar") o.covar = optarg;
    else if (n == "gridL") o.gridL = optarg;
    else if (n == "gridS") o.gridS = optarg;
    else if (n == "bfs") o.bfs = optarg;
    else if (n == "error") o.error = optarg;
    else if (n == "fiterr") o.fiterr = atof(optarg);
    else if (n == "nperm") o.nb_permutations = atol(optarg);
    else if (n == "seed") o.seed = atol(optarg);
    else if (n == "trick") o.trick = atoi(optarg);
    else if (n == "tricut") o.tricut = atol(optarg);
    else if (n == "permsep") o.perm_sep = atoi(optarg);
    else if (n == "pbf") o.pbf = optarg;
    else if (n == "maxbf") o.maxbf = true;
    else if (n == "thread") o.nb_threads = atoi(optarg);
    else if (n == "snp") o.snp = optarg;
    else if (n == "sbgrp") split(optarg, "+", o.sbgrp);
    else if (n == "wrtsize") o.wrtsize = atoi(optarg);
    else if (n == "device") o.device = atoi(optarg);
    else if (n == "gpus") {
      o.gpus = atoi(optarg);
      if (o.gpus < 1 || o.gpus > 64) die_usage(argc, argv, "--gpus should be between 1 and 64");
    }
    else if (n == "shard") {
      if (sscanf(optarg, "%d/%d", &o.shard_k, &o.shard_n) != 2 || o.shard_n < 1 || o.shard_k < 0 || o.shard_k >= o.shard_n)
        die_usage(argc, argv, "--shard should be k/N with 0 <= k < N");
    }
  }
  // validation: same conditions and messages as eqtlbma_bf.cpp:463-692
  if (o.inss.empty()) {
    if (o.geno.empty()) die_usage(argc, argv, "missing compulsory option --geno");
    if (!file_exists(o.geno)) die_usage(argc, argv, "can't find " + o.geno);
  } else { // eqtlbma_bf.cpp:518-537
    if (!file_exists(o.inss)) die_usage(argc, argv, "can't find " + o.inss);
    if (o.analys != "join") die_usage(argc, argv, "--inss requires --analys join");
    if (o.error != "uvlr") die_usage(argc, argv, "--inss requires --error uvlr");
    if (o.nb_permutations > 0) die_usage(argc, argv, "--inss cannot be combined with --nperm (permutations need the raw data)");
    if (o.gpus > 1 || o.shard_n > 1) die_usage(argc, argv, "--inss runs on one GPU");
  }
  if (!o.scoord.empty() && !file_exists(o.scoord)) die_usage(argc, argv, "can't find " + o.scoord);
  if (o.inss.empty()) {
    if (o.exp.empty()) die_usage(argc, argv, "missing compulsory option --exp");
    if (!file_exists(o.exp)) die_usage(argc, argv, "can't find " + o.exp);
    if (o.gcoord.empty()) die_usage(argc, argv, "missing compulsory option --gcoord");
    if (!file_exists(o.gcoord)) die_usage(argc, argv, "can't find " + o.gcoord);
  }
  if (o.anchor != "TSS" && o.anchor != "TSS+TES") die_usage(argc, argv, "--anchor should be TSS or TSS+TES");
  if (o.out.empty()) die_usage(argc, argv, "missing compulsory option --out");
  if (o.wrtsize < 1) die_usage(argc, argv, "--wrtsize should be greater than 1");
  if (o.lik != "normal") die_usage(argc, argv, "--lik " + o.lik + " is not supported by the B200 front-end (out of scope)");
  if (o.analys.empty()) die_usage(argc, argv, "missing compulsory option --analys");
  if (o.analys != "sep" && o.analys != "join") die_usage(argc, argv, "--analys " + o.analys + " is not valid");
  if (o.analys != "join" && o.gridL.empty()) // the reference's inverted test (eqtlbma_bf.cpp:591-596), kept
    die_usage(argc, argv, "missing compulsory option --gridL with --analys join");
  if (!o.gridL.empty() && !file_exists(o.gridL)) die_usage(argc, argv, "can't find " + o.gridL);
  if (o.analys == "join" && (o.bfs == "sin" || o.bfs == "all") && o.gridS.empty())
    die_usage(argc, argv, "--gridS is required with --analys join and --bfs " + o.bfs);
  if (o.bfs != "gen" && o.bfs != "sin" && o.bfs != "all") die_usage(argc, argv, "--bfs " + o.bfs + " is not valid");
  if (o.error != "uvlr" && o.error != "mvlr" && o.error != "hybrid") die_usage(argc, argv, "--error " + o.error + " is not valid");
  if (o.analys == "join" && o.error == "mvlr")
    cerr << "WARNING: summary statistics per subgroup won't be saved with --error mvlr" << endl;
  if (o.trick != 0 && o.trick != 1 && o.trick != 2) die_usage(argc, argv, "--trick is not valid");
  if (o.trick != 0 && o.tricut > o.nb_permutations) die_usage(argc, argv, "--tricut is larger than --nperm");
  if (o.perm_sep != 0 && o.perm_sep != 1 && o.perm_sep != 2) die_usage(argc, argv, "--permsep is not valid");
  if (o.analys == "sep" && o.nb_permutations > 0 && o.perm_sep != 1 && o.perm_sep != 2)
    die_usage(argc, argv, "if --type sep --nperm > 0, --permsep should be '1' or '2'");
  if (o.pbf != "none" && o.pbf != "gen" && o.pbf != "gen-sin" && o.pbf != "all")
    die_usage(argc, argv, "--pbf " + o.pbf + " is unvalid");
  if (o.analys == "join" && o.nb_permutations > 0 && o.pbf == "none")
    die_usage(argc, argv, "if --analys join --nperm > 0, --pbf should be different than 'none'");
  if (o.analys == "join" && o.nb_permutations > 0 && o.bfs == "gen" && o.pbf != "gen")
    die_usage(argc, argv, "if --analys join --bfs gen --nperm > 0, --pbf should be 'gen'");
  if (o.analys == "join" && o.nb_permutations > 0 && o.bfs == "sin" && o.pbf == "all")
    die_usage(argc, argv, "if --analys join --bfs sin --nperm > 0, --pbf should be 'gen' or 'gen-sin'");
  if (o.nb_threads <= 0) die_usage(argc, argv, "--thread is invalid");
  if (!o.snp.empty() && !file_exists(o.snp)) die_usage(argc, argv, "can't find " + o.snp);
  if (o.seed == string::npos) { // utils::getSeed (utils_math.cpp:52-57)
    timeval t1;
    gettimeofday(&t1, NULL);
    o.seed = (size_t)t1.tv_usec * t1.tv_sec;
  }
}

// ------------------------------------------------------------------ loaders (data_loader.cpp)
map<string, string> load_two_column_file(const string &file, int verbose)
{
  map<string, string> m;
  if (file.empty()) return m;
  GzReader r(file);
  if (verbose > 0) cout << "load file " << file << " ..." << endl;
  string line;
  vector<string> tok;
  size_t nb = 0;
  while (r.getline(line)) {
    ++nb;
    split(line, " \t,", tok);
    if (tok.size() != 2) {
      cerr << "ERROR: file " << file << " should have only two columns at line " << nb << endl;
      eqb_exit(EXIT_FAILURE);
    }
    if (tok[0][0] == '#') continue;
    if (m.find(tok[0]) == m.end()) m.insert(make_pair(tok[0], tok[1]));
  }
  if (verbose > 0) cout << "items loaded: " << m.size() << endl;
  return m;
}

vector<string> header_samples(const string &file)
{
  GzReader r(file);
  string line;
  if (!r.getline(line) || line.empty()) {
    cerr << "ERROR: problem with the header of file " << file << endl;
    eqb_exit(EXIT_FAILURE);
  }
  vector<string> tok;
  split(line, " \t", tok);
  if (!tok.empty() && (tok[0] == "Id" || tok[0] == "id" || tok[0] == "ID")) tok.erase(tok.begin());
  set<string> uniq(tok.begin(), tok.end());
  if (uniq.size() != tok.size()) {
    cerr << "ERROR: file " << file << " has redundant samples in its header";
    eqb_exit(EXIT_FAILURE);
  }
  return tok;
}

enum GenoFormat { FMT_DOSE, FMT_VCF, FMT_IMPUTE };

// samples of a genotype file + its format (loadSamplesFromGenotypes, data_loader.cpp:232-332)
vector<string> genotype_samples(const string &file, GenoFormat &fmt)
{
  GzReader r(file);
  string line;
  vector<string> tok;
  if (!r.getline(line) || line.empty()) {
    cerr << "ERROR: problem with the header of file " << file << endl;
    eqb_exit(EXIT_FAILURE);
  }
  if (line.find("##fileformat=VCF") != string::npos) {
    fmt = FMT_VCF;
    while (r.getline(line)) {
      if (line.find("#CHROM") == string::npos) continue;
      split(line, " \t", tok);
      return vector<string>(tok.begin() + 9, tok.end());
    }
    return vector<string>();
  }
  split(line, " \t", tok);
  if (tok.size() >= 5 && tok[0] == "chr" && (tok[1] == "name" || tok[1] == "id") && tok[2] == "coord" &&
      tok[3] == "a1" && tok[4] == "a2") {
    fmt = FMT_IMPUTE;
    if ((tok.size() - 5) % 3 != 0) {
      cerr << "ERROR: the header of IMPUTE file " << file << " is badly formatted" << endl;
      eqb_exit(EXIT_FAILURE);
    }
    vector<string> s;
    for (size_t i = 5; i < tok.size(); i += 3) {
      vector<string> t2;
      split(tok[i], "_a", t2);
      s.push_back(t2[0]);
    }
    return s;
  }
  fmt = FMT_DOSE;
  if (tok[0] == "Id" || tok[0] == "id" || tok[0] == "ID") tok.erase(tok.begin());
  return tok;
}

struct SnpRec {
  string name, chr;
  long pos;
  map<string, vector<double> > geno; // subgroup -> dosages (file column order)
  map<string, double> maf;
};

struct GeneRec {
  string name, chr;
  long start, end; // 1-based start (BED start + 1)
  map<string, vector<double> > exp;
};

// Snp::AddSubgroupFrom{Dose,Vcf,Impute}Line (snp.cpp:88-185): dosages + folded MAF, NaN MAF = missing
void parse_genotypes(GenoFormat fmt, const vector<string> &tok, size_t first, size_t idx_gt, vector<double> &g, double &maf)
{
  maf = 0.0;
  if (fmt == FMT_DOSE) {
    const size_t n = tok.size() - first;
    g.assign(n, kNaN);
    for (size_t i = 0; i < n; ++i) {
      const string &t = tok[first + i];
      if (is_na(t))
        maf = kNaN;
      else {
        g[i] = atof(t.c_str());
        if (maf == maf) maf += g[i];
      }
    }
    if (maf == maf) maf /= (2 * n);
  } else if (fmt == FMT_VCF) {
    const size_t n = tok.size() - first;
    g.assign(n, kNaN);
    vector<string> t2, t3;
    for (size_t i = 0; i < n; ++i) {
      split(tok[first + i], ":", t2);
      if (t2[idx_gt].find(".") != string::npos)
        maf = kNaN;
      else {
        split(t2[idx_gt], "|/", t3);
        g[i] = 0;
        if (t3[0] == "1") g[i] += 1;
        if (t3[1] == "1") g[i] += 1;
        if (maf == maf) maf += g[i];
      }
    }
    if (maf == maf) maf /= (2 * n);
  } else {
    const size_t n = (tok.size() - first) / 3;
    g.assign(n, kNaN);
    for (size_t i = 0; i < n; ++i) {
      const double AA = atof(tok[first + 3 * i].c_str()), AB = atof(tok[first + 3 * i + 1].c_str()),
                   BB = atof(tok[first + 3 * i + 2].c_str());
      if (AA == 0 && AB == 0 && BB == 0)
        maf = kNaN;
      else {
        g[i] = 0 * AA + 1 * AB + 2 * BB;
        if (maf == maf) maf += g[i];
      }
    }
    if (maf == maf) maf /= (2 * n);
  }
  if (maf == maf) maf = (maf <= 0.5 ? maf : 1 - maf);
}

struct Loaded {
  vector<string> subgroups, samples;
  map<string, string> genofile, expfile, covfile;
  map<string, vector<string> > geno_samples, exp_samples, cov_samples;
  map<string, string> loaded_from; // genotype file whose content each subgroup ends up with
  map<string, GeneRec> genes;
  map<string, SnpRec> snps;
  map<string, map<string, vector<double> > > covars; // subgroup -> name-sorted covariates
  vector<double> phi2L, oma2L, phi2S, oma2S;
};

void load_grid(const string &file, vector<double> &phi2, vector<double> &oma2, int verbose)
{
  if (file.empty()) return;
  if (verbose > 0) cout << "load grid in " << file << " ..." << endl << flush;
  GzReader r(file);
  string line;
  vector<string> tok;
  while (r.getline(line)) {
    split(line, " \t", tok);
    if (tok.size() != 2) {
      cerr << "ERROR: format of file " << file << " should be phi2<space/tab>oma2" << endl;
      eqb_exit(1);
    }
    phi2.push_back(atof(tok[0].c_str()));
    oma2.push_back(atof(tok[1].c_str()));
  }
  if (verbose > 0) cout << "grid size: " << phi2.size() << endl;
}

void load_all(const Options &o, Loaded &d)
{
  const int verbose = o.verbose;
  // loadListsGenoExplevelAndCovarFiles (data_loader.cpp:81-165)
  d.expfile = load_two_column_file(o.exp, verbose);
  for (map<string, string>::iterator it = d.expfile.begin(); it != d.expfile.end();)
    if (!o.sbgrp.empty() && find(o.sbgrp.begin(), o.sbgrp.end(), it->first) == o.sbgrp.end())
      d.expfile.erase(it++);
    else
      ++it;
  d.genofile = load_two_column_file(o.geno, verbose);
  for (map<string, string>::iterator it = d.genofile.begin(); it != d.genofile.end();)
    if (d.expfile.find(it->first) == d.expfile.end()) d.genofile.erase(it++);
    else ++it;
  for (map<string, string>::iterator it = d.expfile.begin(); it != d.expfile.end();)
    if (d.genofile.find(it->first) == d.genofile.end()) d.expfile.erase(it++);
    else ++it;
  if (o.error != "uvlr")
    for (map<string, string>::iterator it = d.genofile.begin(); it != d.genofile.end(); ++it)
      if (it->second != d.genofile.begin()->second) {
        cerr << "ERROR: --error mvlr/hybrid requires the same genotypes in a single file for all subgroups" << endl;
        eqb_exit(EXIT_FAILURE);
      }
  for (map<string, string>::iterator it = d.expfile.begin(); it != d.expfile.end(); ++it) d.subgroups.push_back(it->first);
  d.covfile = load_two_column_file(o.covar, verbose);
  for (map<string, string>::iterator it = d.covfile.begin(); it != d.covfile.end();)
    if (find(d.subgroups.begin(), d.subgroups.end(), it->first) == d.subgroups.end()) d.covfile.erase(it++);
    else ++it;
  if (verbose > 0) {
    cout << "analyze " << d.subgroups.size() << " subgroup" << (d.subgroups.size() > 1 ? "s" : "") << " (identifier):" << endl;
    for (size_t s = 0; s < d.subgroups.size(); ++s) cout << d.subgroups[s] << " (" << s + 1 << ")" << endl;
  }
  if (d.subgroups.empty()) return;

  // loadSamples (data_loader.cpp:339-392): sorted union of expression and genotype samples
  if (verbose > 0) cout << "load samples ..." << endl << flush;
  set<string> all;
  map<string, GenoFormat> gfmt;
  for (size_t s = 0; s < d.subgroups.size(); ++s) {
    const string &sg = d.subgroups[s];
    d.exp_samples[sg] = header_samples(d.expfile[sg]);
    all.insert(d.exp_samples[sg].begin(), d.exp_samples[sg].end());
    GenoFormat f;
    d.geno_samples[sg] = genotype_samples(d.genofile[sg], f);
    gfmt[sg] = f;
    all.insert(d.geno_samples[sg].begin(), d.geno_samples[sg].end());
  }
  d.samples.assign(all.begin(), all.end()); // std::set order = std::sort order of the reference
  if (verbose > 0) cout << "total nb of samples: " << d.samples.size() << endl << flush;
  for (map<string, string>::iterator it = d.covfile.begin(); it != d.covfile.end(); ++it) {
    d.cov_samples[it->first] = header_samples(it->second);
    for (size_t i = 0; i < d.cov_samples[it->first].size(); ++i)
      if (all.find(d.cov_samples[it->first][i]) == all.end()) {
        cerr << "ERROR: sample " << d.cov_samples[it->first][i]
             << " has covariates but neither expression levels nor genotypes" << endl;
        eqb_exit(EXIT_FAILURE);
      }
  }

  // loadCovariates (data_loader.cpp:1097-1159): name-sorted, no missing value
  for (map<string, string>::iterator it = d.covfile.begin(); it != d.covfile.end(); ++it) {
    GzReader r(it->second);
    string line;
    vector<string> tok;
    r.getline(line);
    const size_t ns = d.cov_samples[it->first].size();
    size_t nb = 1;
    while (r.getline(line)) {
      ++nb;
      split(line, " \t", tok);
      if (tok.size() != ns + 1) {
        cerr << "ERROR: not enough columns on line " << nb << " of file " << it->second << " (" << tok.size()
             << " != " << ns + 1 << ")" << endl;
        eqb_exit(EXIT_FAILURE);
      }
      if (d.covars[it->first].find(tok[0]) != d.covars[it->first].end()) continue;
      vector<double> v(ns);
      for (size_t i = 0; i < ns; ++i) {
        if (is_na(tok[i + 1])) {
          cerr << "ERROR: no missing value allowed, see covariate " << tok[0] << " in subgroup " << it->first << endl;
          eqb_exit(EXIT_FAILURE);
        }
        v[i] = atof(tok[i + 1].c_str());
      }
      d.covars[it->first][tok[0]] = v;
    }
  }

  // loadGeneInfo (data_loader.cpp:396-435)
  if (verbose > 0) cout << "load gene coordinates ..." << endl << flush;
  set<string> gene_chrs;
  {
    GzReader r(o.gcoord);
    string line;
    vector<string> tok;
    while (r.getline(line)) {
      split(line, " \t", tok);
      if (tok.size() < 4) continue;
      if (d.genes.find(tok[3]) != d.genes.end()) continue;
      if (tok[1] == tok[2]) {
        cerr << "ERROR: start and end coordinates of " << tok[3] << " should be different (at least 1 bp)" << endl;
        eqb_exit(1);
      }
      GeneRec g;
      g.name = tok[3];
      g.chr = tok[0];
      g.start = atol(tok[1].c_str()) + 1;
      g.end = atol(tok[2].c_str());
      d.genes[g.name] = g;
      gene_chrs.insert(g.chr);
    }
  }
  if (verbose > 0) cout << "total nb of genes with coordinates: " << d.genes.size() << endl;

  // loadExplevels (data_loader.cpp:437-527)
  if (verbose > 0) cout << "load gene expression levels ..." << endl << flush;
  for (size_t s = 0; s < d.subgroups.size(); ++s) {
    const string &sg = d.subgroups[s];
    GzReader r(d.expfile[sg]);
    string line;
    vector<string> tok;
    r.getline(line);
    const size_t ns = d.exp_samples[sg].size();
    size_t nb = 1, kept = 0;
    vector<Span> sp;
    while (r.getline(line)) {
      ++nb;
      split_spans(line, sp);
      if (sp.size() != ns + 1) {
        cerr << "ERROR: not enough columns on line " << nb << " of file " << d.expfile[sg] << " (" << sp.size()
             << " != " << ns + 1 << ")" << endl;
        eqb_exit(EXIT_FAILURE);
      }
      map<string, GeneRec>::iterator g = d.genes.find(sp[0].str());
      if (g == d.genes.end()) continue;
      if (g->second.exp.find(sg) != g->second.exp.end()) continue; // map::insert keeps the first
      vector<double> v(ns, kNaN);
      for (size_t i = 0; i < ns; ++i)
        if (!is_na(sp[i + 1])) v[i] = fast_atof(sp[i + 1]);
      g->second.exp[sg] = v;
      ++kept;
    }
    if (verbose > 0) cout << sg << " (" << d.expfile[sg] << "): " << (nb - 1) << " genes (to keep: " << kept << ")" << endl << flush;
  }
  for (map<string, GeneRec>::iterator it = d.genes.begin(); it != d.genes.end();) {
    bool any = false;
    for (map<string, vector<double> >::iterator e = it->second.exp.begin(); e != it->second.exp.end(); ++e)
      if (!e->second.empty()) any = true;
    if (!any) d.genes.erase(it++);
    else ++it;
  }
  if (verbose > 0) cout << "total nb of genes to analyze: " << d.genes.size() << endl;
  if (d.genes.empty()) return;

  // --snp
  set<string> snps_to_keep;
  if (!o.snp.empty()) {
    GzReader r(o.snp);
    string line;
    vector<string> tok;
    while (r.getline(line)) {
      split(line, " \t,", tok);
      if (tok.size() != 1) {
        cerr << "ERROR: file " << o.snp << " should have only one column" << endl;
        eqb_exit(EXIT_FAILURE);
      }
      if (tok[0][0] == '#') continue;
      snps_to_keep.insert(tok[0]);
    }
  }

  // SNP coordinates (--scoord: loadSnpInfo, data_loader.cpp:845-876) and genotypes
  // (loadGenos :878-1010 / loadGenosAndSnpInfo :717-843)
  const bool custom = !o.scoord.empty();
  if (custom) {
    if (verbose > 0) cout << "load SNP coordinates (unindexed BED file) ..." << endl << flush;
    GzReader r(o.scoord);
    string line;
    vector<string> tok;
    while (r.getline(line)) {
      split(line, " \t", tok);
      if (tok.size() < 4) continue;
      if (!snps_to_keep.empty() && snps_to_keep.find(tok[3]) == snps_to_keep.end()) continue;
      if (d.snps.find(tok[3]) != d.snps.end()) continue;
      if (tok[1] == tok[2]) {
        cerr << "ERROR: start and end coordinates of " << tok[3] << " should be different (at least 1 bp)" << endl;
        eqb_exit(1);
      }
      SnpRec s;
      s.name = tok[3];
      s.chr = tok[0];
      s.pos = atol(tok[2].c_str());
      d.snps[s.name] = s;
    }
    if (verbose > 0) cout << "total nb of SNPs with coordinates: " << d.snps.size() << endl;
    if (d.snps.empty()) return;
  }
  if (verbose > 0) cout << "load genotypes ..." << endl << flush;
  bool same_files = false;
  const string first_sg = d.genofile.begin()->first;
  for (map<string, string>::iterator it = d.genofile.begin(); it != d.genofile.end(); ++it) {
    if (it != d.genofile.begin() && it->second == d.genofile.begin()->second) {
      same_files = true;
      break; // the reference stops loading at the first repeat of the first file (data_loader.cpp:733-740)
    }
    const string &sg = it->first;
    d.loaded_from[sg] = it->second;
    GenoFormat fmt = gfmt[sg];
    if (custom && fmt != FMT_DOSE) {
      cerr << "ERROR: don't use --scoord if genotypes in IMPUTE or VCF format" << endl;
      eqb_exit(1);
    }
    if (!custom && fmt == FMT_DOSE) {
      cerr << "ERROR: file " << it->second << " seems to be in the custom format but --scoord is missing" << endl;
      eqb_exit(EXIT_FAILURE);
    }
    GzReader r(it->second);
    string line;
    vector<string> tok, t2;
    r.getline(line);
    if (fmt == FMT_VCF)
      while (r.getline(line))
        if (line.find("#CHROM") != string::npos) break;
    const size_t ns = d.geno_samples[sg].size();
    size_t nb = 1, kept = 0;
    vector<Span> sp;
    while (r.getline(line)) {
      ++nb;
      string name, chr, pos;
      size_t first = 1, idx_gt = 0;
      if (fmt == FMT_DOSE) {
        // custom format: in-place tokens, decimal fast path (Snp::AddSubgroupFromDoseLine, snp.cpp:88-113)
        split_spans(line, sp);
        if (sp.size() != ns + 1) {
          cerr << "ERROR: not enough columns on line " << nb << " of file " << it->second << " (" << sp.size()
               << " != " << ns + 1 << ")" << endl;
          eqb_exit(EXIT_FAILURE);
        }
        name = sp[0].str();
        map<string, SnpRec>::iterator si = d.snps.find(name);
        if (si == d.snps.end()) continue;
        bool has_NA = false;
        for (size_t i = 0; i < sp.size() && !has_NA; ++i)
          has_NA = sp[i].n == 2 && sp[i].p[0] == 'N' && sp[i].p[1] == 'A'; // data_loader.cpp:937-938
        if (has_NA) continue;
        SnpRec &sr = si->second;
        vector<double> &g = sr.geno[sg];
        if (!g.empty()) {
          cerr << "ERROR: SNP " << name << " is duplicated in file " << it->second << endl;
          eqb_exit(EXIT_FAILURE);
        }
        g.assign(ns, kNaN);
        double maf = 0.0;
        for (size_t i = 0; i < ns; ++i) {
          if (is_na(sp[i + 1]))
            maf = kNaN;
          else {
            g[i] = fast_atof(sp[i + 1]);
            if (maf == maf) maf += g[i];
          }
        }
        if (maf == maf) maf /= (2 * ns);
        if (maf == maf) maf = (maf <= 0.5 ? maf : 1 - maf);
        sr.maf[sg] = maf;
        ++kept;
        continue;
      }
      split(line, " \t", tok);
      if (false) {
      } else if (fmt == FMT_VCF) {
        if (tok.size() != ns + 9) {
          cerr << "ERROR: not enough columns on line " << nb << " of file " << it->second << endl;
          eqb_exit(EXIT_FAILURE);
        }
        if (tok[8].find("GT") == string::npos) {
          cerr << "ERROR: missing GT in 9-th field on line " << nb << " of file " << it->second << endl;
          eqb_exit(EXIT_FAILURE);
        }
        chr = tok[0];
        pos = tok[1];
        name = tok[2];
        first = 9;
        if (gene_chrs.find(chr) == gene_chrs.end()) continue;
        if (!snps_to_keep.empty() && snps_to_keep.find(name) == snps_to_keep.end()) continue;
        split(tok[8], ":", t2);
        while (idx_gt < t2.size() && t2[idx_gt] != "GT") ++idx_gt;
      } else {
        if (tok.size() != 3 * ns + 5) {
          cerr << "ERROR: not enough columns on line " << nb << " of file " << it->second << endl;
          eqb_exit(EXIT_FAILURE);
        }
        chr = tok[0];
        name = tok[1];
        pos = tok[2];
        first = 5;
        if (gene_chrs.find(chr) == gene_chrs.end()) continue;
        if (!snps_to_keep.empty() && snps_to_keep.find(name) == snps_to_keep.end()) continue;
      }
      if (fmt != FMT_DOSE && d.snps.find(name) == d.snps.end()) {
        SnpRec s;
        s.name = name;
        s.chr = chr;
        s.pos = atol(pos.c_str());
        d.snps[name] = s;
      }
      SnpRec &sr = d.snps[name];
      if (sr.geno.find(sg) != sr.geno.end() && !sr.geno[sg].empty()) {
        cerr << "ERROR: SNP " << name << " is duplicated in file " << it->second << endl;
        eqb_exit(EXIT_FAILURE);
      }
      vector<double> g;
      double maf;
      parse_genotypes(fmt, tok, first, idx_gt, g, maf);
      sr.geno[sg] = g;
      sr.maf[sg] = maf;
      ++kept;
    }
    if (verbose > 0) cout << sg << " (" << it->second << "): " << (nb - 1) << " SNPs (to keep: " << kept << ")" << endl << flush;
  }
  // discard SNP x subgroup with missing values, then MAF filter (snp.cpp:233-253), then SNPs without any subgroup
  if (verbose > 0) cout << "discard SNPs with missing values ..." << endl << flush;
  for (map<string, SnpRec>::iterator it = d.snps.begin(); it != d.snps.end();) {
    SnpRec &sr = it->second;
    for (map<string, double>::iterator m = sr.maf.begin(); m != sr.maf.end();) {
      const bool drop = (m->second != m->second) || (o.min_maf > 0 && m->second < (double)o.min_maf);
      if (drop) {
        sr.geno.erase(m->first);
        sr.maf.erase(m++);
      } else
        ++m;
    }
    bool any = false;
    for (map<string, vector<double> >::iterator g = sr.geno.begin(); g != sr.geno.end(); ++g)
      if (!g->second.empty()) any = true;
    if (!any) d.snps.erase(it++);
    else ++it;
  }
  if (same_files) {
    // Snp::DuplicateGenotypesFromFirstSubgroup for every other subgroup (data_loader.cpp:691-715);
    // map::insert semantics: a subgroup that already has genotypes keeps them
    for (map<string, SnpRec>::iterator it = d.snps.begin(); it != d.snps.end(); ++it)
      for (map<string, string>::iterator sgi = d.genofile.begin(); sgi != d.genofile.end(); ++sgi) {
        if (sgi->first == first_sg) continue;
        if (it->second.geno.find(sgi->first) != it->second.geno.end()) continue;
        if (it->second.geno.find(first_sg) == it->second.geno.end()) {
          it->second.geno[sgi->first] = vector<double>(); // operator[] of the reference creates an empty entry
          it->second.maf[sgi->first] = kNaN;
        } else {
          it->second.geno[sgi->first] = it->second.geno[first_sg];
          it->second.maf[sgi->first] = it->second.maf[first_sg];
        }
      }
  }
  // subgroups that were not loaded themselves carry the first file's genotypes (and sample list)
  for (map<string, string>::iterator sgi = d.genofile.begin(); sgi != d.genofile.end(); ++sgi)
    if (d.loaded_from.find(sgi->first) == d.loaded_from.end()) {
      if (sgi->second != d.genofile.begin()->second)
        // (the reference stops reading genotype files at the first repeat of the first one, data_loader.cpp:733-740, and
        // pairs the duplicated vectors with this file's own sample list; the listed file is ignored by both programs)
        cerr << "WARNING: genotype file " << sgi->second << " of subgroup " << sgi->first << " is not read: the list repeats "
             << d.genofile.begin()->second << " before it, and its genotypes and samples are used instead" << endl;
      d.loaded_from[sgi->first] = d.genofile.begin()->second;
      d.geno_samples[sgi->first] = d.geno_samples[first_sg];
    }
  if (verbose > 0) cout << "total nb of SNPs to analyze: " << d.snps.size() << endl;
  load_grid(o.gridL, d.phi2L, d.oma2L, verbose);
  load_grid(o.gridS, d.phi2S, d.oma2S, verbose);
}

// wall-clock phases of a run, printed with -v 2 and above (extension; the reference only reports the total)
struct PhaseClock {
  double t0, last;
  vector<pair<string, double> > ph;
  static double now()
  {
    timeval tv;
    gettimeofday(&tv, NULL);
    return (double)tv.tv_sec + 1e-6 * (double)tv.tv_usec;
  }
  PhaseClock() : t0(now()), last(t0) {}
  void mark(const string &name)
  {
    const double t = now();
    for (size_t i = 0; i < ph.size(); ++i)
      if (ph[i].first == name) {
        ph[i].second += t - last;
        last = t;
        return;
      }
    ph.push_back(make_pair(name, t - last));
    last = t;
  }
  void print() const
  {
    cout << "phases (wall clock, s):";
    for (size_t i = 0; i < ph.size(); ++i) cout << " " << ph[i].first << "=" << ph[i].second;
    cout << " total=" << now() - t0 << endl;
  }
};

void check(eqb_ctx *ctx, int rc, const char *what)
{
  if (rc != 0) {
    cerr << "ERROR: " << what << ": " << eqb_last_error(ctx) << endl;
    eqb_exit(EXIT_FAILURE);
  }
}

vector<string> config_names(int S, const string &bfs)
{
  vector<string> names;
  if (bfs == "gen") return names;
  for (int k = 1; k <= S; ++k) {
    vector<int> d(k);
    for (int i = 0; i < k; ++i) d[i] = i;
    while (true) {
      stringstream ss;
      ss << d[0] + 1;
      for (int i = 1; i < k; ++i) ss << "-" << d[i] + 1;
      names.push_back(ss.str());
      int i = k - 1;
      while (i > 0 && d[i] == S - k + i) --i;
      if (i == 0 && d[i] == S - k) break;
      ++d[i];
      for (; i < k - 1; ++i) d[i + 1] = d[i] + 1;
    }
    if (bfs == "sin") break;
  }
  return names;
}

// --gpus N: the reference spreads gene batches over OS processes with scripts/eqtlbma_bf_parallel.bash:248-345 and
// merges with `zcat | sed 1d` (doc/manual_eqtlbma.texi:1211-1218).  Here: one child per GPU with --shard k/N (contiguous
// cost-balanced ranges of whole write-groups, eqb_partition_by_cost), each writing <out>.shard<k>_*; shard 0 writes the
// header members, the others none, so the byte-wise concatenation in shard order IS the output of a single run
// (gzip members concatenate; same lines, same order).  Returns the exit status for main(), or -1 in a child.
int launch_shards(int argc, char **argv, Options &o, time_t t_start)
{
  const string out = o.out;
  vector<pid_t> pids(o.gpus);
  fflush(stdout);
  fflush(stderr);
  for (int k = 0; k < o.gpus; ++k) {
    const pid_t pid = fork(); // (before any CUDA call: every child creates its own context on its own device)
    if (pid < 0) {
      cerr << "ERROR: fork failed" << endl;
      return EXIT_FAILURE;
    }
    if (pid == 0) {
      o.shard_k = k;
      o.shard_n = o.gpus;
      o.device = k;
      o.out = out + ".shard" + to_string(k);
      o.shard_child = true;
      o.no_header = k > 0;
      if (k > 0) o.verbose = 0;
      return -1;
    }
    pids[k] = pid;
  }
  bool ok = true;
  for (int k = 0; k < o.gpus; ++k) {
    int status = 0;
    if (waitpid(pids[k], &status, 0) < 0 || !WIFEXITED(status) || WEXITSTATUS(status) != 0) ok = false;
  }
  if (!ok) {
    cerr << "ERROR: a shard process failed" << endl;
    return EXIT_FAILURE;
  }
  // merge: every file <out>.shard0<suffix> with its counterparts of the other shards
  string dir = ".", base = out;
  const size_t slash = out.find_last_of('/');
  if (slash != string::npos) {
    dir = out.substr(0, slash);
    base = out.substr(slash + 1);
  }
  const string pre0 = base + ".shard0";
  vector<string> suffixes;
  if (DIR *dp = opendir(dir.c_str())) {
    while (struct dirent *e = readdir(dp)) {
      const string fn = e->d_name;
      if (fn.compare(0, pre0.size(), pre0) == 0 && fn.size() > pre0.size() && fn[pre0.size()] == '_')
        suffixes.push_back(fn.substr(pre0.size()));
    }
    closedir(dp);
  }
  sort(suffixes.begin(), suffixes.end());
  vector<char> buf(1 << 22);
  for (const string &suf : suffixes) {
    FILE *fo = fopen((out + suf).c_str(), "wb");
    if (fo == NULL) {
      cerr << "ERROR: can't open file " << out + suf << " with mode wb" << endl;
      return EXIT_FAILURE;
    }
    for (int k = 0; k < o.gpus; ++k) {
      const string part = out + ".shard" + to_string(k) + suf;
      FILE *fi = fopen(part.c_str(), "rb");
      if (fi == NULL) continue;
      size_t n;
      while ((n = fread(buf.data(), 1, buf.size(), fi)) > 0)
        if (fwrite(buf.data(), 1, n, fo) != n) {
          cerr << "ERROR: can't write to file " << out + suf << endl;
          return EXIT_FAILURE;
        }
      fclose(fi);
      remove(part.c_str());
    }
    fclose(fo);
  }
  size_t pairs = 0, genes = 0;
  for (int k = 0; k < o.gpus; ++k) {
    const string cf = out + ".shard" + to_string(k) + ".count";
    if (FILE *f = fopen(cf.c_str(), "r")) {
      size_t a = 0, b = 0;
      if (fscanf(f, "%zu %zu", &a, &b) == 2) {
        pairs += a;
        genes += b;
      }
      fclose(f);
      remove(cf.c_str());
    }
  }
  if (o.verbose > 0) {
    cout << "nb of analyzed gene-SNP pairs: " << pairs << " (" << genes << " genes)" << endl;
    time_t t_end;
    time(&t_end);
    cout << "END " << argv[0] << " " << ctime(&t_end) << "elapsed -> " << difftime(t_end, t_start) << " sec" << endl;
  }
  (void)argc;
  return EXIT_SUCCESS;
}

} // namespace

// --inss: Bayes factors from per-subgroup summary-statistics files (the `_sumstats_<subgroup>.txt.gz` files --outss writes)
// instead of raw data: loadSummaryStats / fillGeneSnpPairsWithSstats (data_loader.cpp:1212-1343), then the write-group loop
// of run() with hasDataNotSstats = false (eqtlbma_bf.cpp:1546-1574).  Genes in name order; the SNPs of a gene in the order
// of their first appearance (subgroups in name order, file order inside a subgroup); a (pair, subgroup) listed twice keeps
// its first line (map::insert).
int run_inss(const Options &o, char **argv, time_t t_start)
{
  map<string, string> files = load_two_column_file(o.inss, o.verbose);
  vector<string> subgroups;
  for (map<string, string>::iterator it = files.begin(); it != files.end(); ++it) subgroups.push_back(it->first);
  const int S = (int)subgroups.size();
  struct PairIn {
    string snp;
    vector<int32_t> n;
    vector<double> sig, beta, se;
  };
  map<string, vector<PairIn> > genes;
  map<string, map<string, size_t> > snp_idx; // gene -> snp -> index in its vector
  static const char *cols[6] = {"gene", "snp", "n", "sigmahat", "betahat.geno", "sebetahat.geno"};
  for (int s = 0; s < S; ++s) {
    if (o.verbose > 0) cout << "load summary statistics for subgroup " << subgroups[s] << " ..." << endl;
    GzReader r(files[subgroups[s]]);
    string line;
    vector<string> tok;
    if (!r.getline(line)) continue;
    split(line, "\t", tok);
    size_t ci[6];
    for (int c = 0; c < 6; ++c) {
      ci[c] = string::npos;
      for (size_t i = 0; i < tok.size(); ++i)
        if (tok[i] == cols[c]) ci[c] = i;
      if (ci[c] == string::npos) {
        cerr << "ERROR: missing " << cols[c] << " in header of " << files[subgroups[s]] << endl;
        eqb_exit(EXIT_FAILURE);
      }
    }
    while (r.getline(line)) {
      split(line, "\t", tok);
      size_t need = 0;
      for (int c = 0; c < 6; ++c) need = max(need, ci[c] + 1);
      if (tok.size() < need) continue;
      vector<PairIn> &v = genes[tok[ci[0]]];
      map<string, size_t> &ix = snp_idx[tok[ci[0]]];
      map<string, size_t>::iterator f = ix.find(tok[ci[1]]);
      size_t j;
      if (f == ix.end()) {
        j = v.size();
        ix[tok[ci[1]]] = j;
        PairIn pi;
        pi.snp = tok[ci[1]];
        pi.n.assign(S, 0);
        pi.sig.assign(S, kNaN);
        pi.beta.assign(S, kNaN);
        pi.se.assign(S, kNaN);
        v.push_back(pi);
      } else
        j = f->second;
      if (v[j].n[s] > 0) continue; // insert() keeps the first
      v[j].n[s] = (int32_t)atol(tok[ci[2]].c_str());
      v[j].sig[s] = atof(tok[ci[3]].c_str());
      v[j].beta[s] = atof(tok[ci[4]].c_str());
      v[j].se[s] = atof(tok[ci[5]].c_str());
    }
  }
  if (genes.empty()) return EXIT_SUCCESS;
  vector<double> phi2L, oma2L, phi2S, oma2S;
  load_grid(o.gridL, phi2L, oma2L, o.verbose);
  load_grid(o.gridS, phi2S, oma2S, o.verbose);
  const int L = (int)phi2L.size(), K = (int)phi2S.size();
  const vector<string> cnames = config_names(S, o.bfs);
  const int64_t C = (int64_t)cnames.size();
  // headers (writeRes(..., "only"))
  const string sep = "\t";
  {
    string h = "gene\tsnp\tconfig";
    for (int i = 0; i < L; ++i) h += "\tl10abf.grid" + to_string(i + 1);
    gz_write(o.out + "_l10abfs_raw.txt.gz", "wb", h + "\n");
    if (o.outw) {
      h = "gene\tsnp\tnb.subgroups\tl10abf.gen\tl10abf.gen.fix\tl10abf.gen.maxh";
      if (o.bfs != "gen") h += "\tl10abf.gen.sin";
      if (o.bfs == "all") h += "\tl10abf.all";
      for (size_t c = 0; c < cnames.size(); ++c) h += "\tl10abf." + cnames[c];
      gz_write(o.out + "_l10abfs_avg-grids.txt.gz", "wb", h + "\n");
    }
  }
  if (o.verbose > 0)
    cout << "test for association between each pair gene-SNP ..." << endl
         << "analysis=" << o.analys << " likelihood=" << o.lik << " error_model=" << o.error << endl
         << flush;
  eqb_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.abi_version = EQB_ABI_VERSION;
  cfg.n_subgroups = S;
  cfg.n_samples_all = 1;
  cfg.analysis = EQB_ANALYSIS_JOIN;
  cfg.bfs = o.bfs == "gen" ? EQB_BFS_GEN : (o.bfs == "sin" ? EQB_BFS_SIN : EQB_BFS_ALL);
  cfg.error_model = EQB_ERROR_UVLR;
  cfg.device = o.device;
  cfg.fiterr = o.fiterr;
  eqb_ctx *ctx = NULL;
  check(ctx, eqb_create(&ctx, &cfg), "eqb_create");
  check(ctx, eqb_set_grids(ctx, phi2L.data(), oma2L.data(), L, phi2S.data(), oma2S.data(), K), "eqb_set_grids");
  // batches of whole genes sized by the raw-ABF budget
  vector<const string *> gname;
  vector<const vector<PairIn> *> gpairs;
  for (map<string, vector<PairIn> >::iterator it = genes.begin(); it != genes.end(); ++it) {
    gname.push_back(&it->first);
    gpairs.push_back(&it->second);
  }
  const int64_t G = (int64_t)gname.size();
  const size_t per_pair = ((size_t)3 * L + (size_t)C * K + 5 + C) * 8 + (size_t)S * 28;
  const size_t budget = (size_t)1 << 30;
  const int nthr = max(1, o.nb_threads);
  static const char *rows[3] = {"gen", "gen-fix", "gen-maxh"};
  size_t nbPairs = 0;
  for (int64_t g0 = 0; g0 < G;) {
    int64_t g1 = g0;
    size_t np = 0;
    while (g1 < G && (g1 == g0 || (np + gpairs[g1]->size()) * per_pair <= budget)) np += gpairs[g1++]->size();
    vector<int64_t> off(g1 - g0 + 1, 0);
    for (int64_t g = g0; g < g1; ++g) off[g - g0 + 1] = off[g - g0] + (int64_t)gpairs[g]->size();
    const int64_t P = off[g1 - g0];
    vector<int32_t> n((size_t)P * S);
    vector<double> sig((size_t)P * S), beta((size_t)P * S), se((size_t)P * S);
    for (int64_t g = g0; g < g1; ++g)
      for (size_t j = 0; j < gpairs[g]->size(); ++j) {
        const PairIn &pi = (*gpairs[g])[j];
        const size_t p = (size_t)off[g - g0] + j;
        for (int s = 0; s < S; ++s) {
          n[p * S + s] = pi.n[s];
          sig[p * S + s] = pi.sig[s];
          beta[p * S + s] = pi.beta[s];
          se[p * S + s] = pi.se[s];
        }
      }
    vector<double> agen((size_t)P * 3 * L), acfg((size_t)P * C * K), aw((size_t)P * (5 + C));
    eqb_results res;
    memset(&res, 0, sizeof(res));
    res.abf_gen = agen.data();
    res.abf_cfg = acfg.data();
    res.abf_w = aw.data();
    check(ctx, eqb_bf_from_sstats(ctx, P, n.data(), sig.data(), beta.data(), se.data(), &res), "eqb_bf_from_sstats");
    parallel_emit(o.out + "_l10abfs_raw.txt.gz", nthr, g0, g1, off, [&](int64_t ga, int64_t gb, string &raw) {
      for (int64_t g = ga; g < gb; ++g)
        for (size_t j = 0; j < gpairs[g]->size(); ++j) {
          const int64_t p = off[g - g0] + (int64_t)j;
          const string &gn = *gname[g], &sn = (*gpairs[g])[j].snp;
          for (int r = 0; r < 3; ++r) {
            raw += gn + sep + sn + sep + rows[r];
            for (int k = 0; k < L; ++k) {
              raw += sep;
              put_sci(raw, agen[(p * 3 + r) * L + k]);
            }
            raw += "\n";
          }
          for (int64_t c = 0; c < C; ++c) {
            raw += gn + sep + sn + sep + cnames[c];
            for (int k = 0; k < L; ++k) { // padded / truncated to |gridL| columns (eqtlbma_bf.cpp:1207-1212)
              raw += sep;
              put_sci(raw, k < K ? acfg[(p * C + c) * K + k] : kNaN);
            }
            raw += "\n";
          }
        }
    });
    if (o.outw)
      parallel_emit(o.out + "_l10abfs_avg-grids.txt.gz", nthr, g0, g1, off, [&](int64_t ga, int64_t gb, string &avg) {
        for (int64_t g = ga; g < gb; ++g)
          for (size_t j = 0; j < gpairs[g]->size(); ++j) {
            const int64_t p = off[g - g0] + (int64_t)j;
            int nsub = 0;
            for (int s = 0; s < S; ++s) nsub += n[p * S + s] > 0 ? 1 : 0;
            avg += *gname[g] + sep + (*gpairs[g])[j].snp + sep + to_string(nsub);
            const double *w = &aw[p * (5 + C)];
            for (int k = 0; k < 3; ++k) {
              avg += sep;
              put_sci(avg, w[k]);
            }
            if (o.bfs != "gen") {
              avg += sep;
              put_sci(avg, w[3]);
            }
            if (o.bfs == "all") {
              avg += sep;
              put_sci(avg, w[4]);
            }
            for (int64_t c = 0; c < C; ++c) {
              avg += sep;
              put_sci(avg, w[5 + c]);
            }
            avg += "\n";
          }
      });
    nbPairs += (size_t)P;
    g0 = g1;
  }
  eqb_destroy(ctx);
  if (o.verbose > 0) {
    cout << "nb of analyzed gene-SNP pairs: " << nbPairs << " (" << G << " genes)" << endl;
    time_t t_end;
    time(&t_end);
    cout << "END " << argv[0] << " " << ctime(&t_end) << "elapsed -> " << difftime(t_end, t_start) << " sec" << endl;
  }
  return EXIT_SUCCESS;
}

int main(int argc, char **argv)
{
  Options o;
  parse_cmdline(argc, argv, o);
  time_t t_start;
  time(&t_start);
  if (o.verbose > 0) {
    cout << "START " << argv[0] << " " << ctime(&t_start) << "version " << EQB_VERSION << " (B200-native hot path)" << endl
         << "cmd-line:";
    for (int i = 0; i < argc; ++i) cout << " " << argv[i];
    cout << endl << flush;
  }
  if (o.gpus > 1 && o.shard_n == 1) {
    const int rc = launch_shards(argc, argv, o, t_start);
    if (rc >= 0) return rc; // launcher (parent); children fall through with their shard options
    const int ndev = eqb_device_count(); // (first CUDA call of the child: after the fork)
    if (ndev > 0) o.device = o.shard_k % ndev; // fewer GPUs than shards: the shards share them
  }
  if (!o.inss.empty()) return run_inss(o, argv, t_start);
  PhaseClock pc;
  // the CUDA context is created while the input files are parsed (after the --gpus fork: one context per shard process)
  std::thread warm([&o] { eqb_warmup(o.device); });
  g_warm = &warm;
  Loaded d;
  load_all(o, d);
  pc.mark("load");
  warm.join();
  g_warm = NULL;
  pc.mark("cuda-start");
  if (d.genes.empty() || d.snps.empty()) return EXIT_SUCCESS;

  const int S = (int)d.subgroups.size(), N = (int)d.samples.size();
  const bool join = o.analys == "join";
  // ---- index spaces
  map<string, int> sample_idx;
  for (int i = 0; i < N; ++i) sample_idx[d.samples[i]] = i;
  vector<string> chr_names;
  {
    set<string> cs;
    for (map<string, SnpRec>::iterator it = d.snps.begin(); it != d.snps.end(); ++it) cs.insert(it->second.chr);
    for (map<string, GeneRec>::iterator it = d.genes.begin(); it != d.genes.end(); ++it) cs.insert(it->second.chr);
    chr_names.assign(cs.begin(), cs.end());
  }
  map<string, int> chr_idx;
  for (size_t c = 0; c < chr_names.size(); ++c) chr_idx[chr_names[c]] = (int)c;
  // SNP order: chromosome, then position (ties: name order, the reference's std::sort is unstable there)
  vector<const SnpRec *> snps;
  for (map<string, SnpRec>::iterator it = d.snps.begin(); it != d.snps.end(); ++it) snps.push_back(&it->second);
  stable_sort(snps.begin(), snps.end(), [&](const SnpRec *a, const SnpRec *b) {
    if (a->chr != b->chr) return chr_idx[a->chr] < chr_idx[b->chr];
    return a->pos < b->pos;
  });
  const int64_t M = (int64_t)snps.size();
  vector<const GeneRec *> genes;
  for (map<string, GeneRec>::iterator it = d.genes.begin(); it != d.genes.end(); ++it) genes.push_back(&it->second);
  const int64_t G = (int64_t)genes.size();

  // ---- context
  eqb_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.abi_version = EQB_ABI_VERSION;
  cfg.n_subgroups = S;
  cfg.n_samples_all = N;
  cfg.analysis = join ? EQB_ANALYSIS_JOIN : EQB_ANALYSIS_SEP;
  cfg.n_snps = M;
  cfg.n_genes = G;
  cfg.bfs = o.bfs == "gen" ? EQB_BFS_GEN : (o.bfs == "sin" ? EQB_BFS_SIN : EQB_BFS_ALL);
  cfg.error_model = o.error == "mvlr" ? EQB_ERROR_MVLR : (o.error == "hybrid" ? EQB_ERROR_HYBRID : EQB_ERROR_UVLR);
  cfg.qnorm = o.qnorm ? 1 : 0;
  cfg.device = o.device;
  cfg.fiterr = o.fiterr;
  eqb_ctx *ctx = NULL;
  pc.mark("index");
  check(ctx, eqb_create(&ctx, &cfg), "eqb_create");
  pc.mark("create");

  // genotype matrices: one per genotype file actually loaded (subgroups that share a file share it)
  map<string, int> path2gid;
  vector<vector<double> > Gmats;
  vector<int> Gcols;
  vector<int> sub_geno_id(S);
  vector<vector<int32_t> > all2geno(S), all2exp(S), all2cov(S);
  vector<vector<uint8_t> > snp_has(S), gene_has(S);
  vector<vector<double> > Ymats(S), Cmats(S);
  for (int s = 0; s < S; ++s) {
    const string &sg = d.subgroups[s];
    const string &path = d.loaded_from[sg];
    const vector<string> &gs = d.geno_samples[sg];
    const int ncols = (int)gs.size();
    int gid;
    if (path2gid.find(path) == path2gid.end()) {
      gid = (int)Gmats.size();
      path2gid[path] = gid;
      Gmats.push_back(vector<double>((size_t)M * ncols, 0.0));
      Gcols.push_back(ncols);
    } else
      gid = path2gid[path];
    snp_has[s].assign(M, 0);
    for (int64_t m = 0; m < M; ++m) {
      map<string, vector<double> >::const_iterator g = snps[m]->geno.find(sg);
      if (g != snps[m]->geno.end() && !g->second.empty() && (int)g->second.size() == Gcols[gid]) {
        snp_has[s][m] = 1;
        memcpy(&Gmats[gid][(size_t)m * ncols], g->second.data(), ncols * sizeof(double));
      }
    }
    sub_geno_id[s] = gid;
    all2geno[s].assign(N, -1);
    for (int i = 0; i < ncols; ++i) all2geno[s][sample_idx[gs[i]]] = i;
    const vector<string> &es = d.exp_samples[sg];
    all2exp[s].assign(N, -1);
    for (size_t i = 0; i < es.size(); ++i) all2exp[s][sample_idx[es[i]]] = (int)i;
    gene_has[s].assign(G, 0);
    Ymats[s].assign((size_t)G * es.size(), kNaN);
    for (int64_t g = 0; g < G; ++g) {
      map<string, vector<double> >::const_iterator e = genes[g]->exp.find(sg);
      if (e != genes[g]->exp.end()) {
        gene_has[s][g] = 1;
        memcpy(&Ymats[s][(size_t)g * es.size()], e->second.data(), es.size() * sizeof(double));
      }
    }
    all2cov[s].assign(N, -1);
    int Q = 0, ncc = 0;
    if (d.covars.find(sg) != d.covars.end()) {
      const vector<string> &cs = d.cov_samples[sg];
      ncc = (int)cs.size();
      for (int i = 0; i < ncc; ++i) all2cov[s][sample_idx[cs[i]]] = i;
      Q = (int)d.covars[sg].size();
      Cmats[s].resize((size_t)Q * ncc);
      int q = 0;
      for (map<string, vector<double> >::iterator c = d.covars[sg].begin(); c != d.covars[sg].end(); ++c, ++q)
        memcpy(&Cmats[s][(size_t)q * ncc], c->second.data(), ncc * sizeof(double));
    }
    eqb_subgroup sub;
    memset(&sub, 0, sizeof(sub));
    sub.geno_id = gid;
    sub.n_exp_cols = (int)es.size();
    sub.n_covariates = Q;
    sub.n_cov_cols = ncc;
    sub.all2geno = all2geno[s].data();
    sub.all2exp = all2exp[s].data();
    sub.all2cov = all2cov[s].data();
    sub.snp_has_geno = snp_has[s].data();
    sub.gene_has_exp = gene_has[s].data();
    sub.Y = Ymats[s].data();
    sub.C = Q ? Cmats[s].data() : NULL;
    check(ctx, eqb_set_subgroup(ctx, s, &sub), "eqb_set_subgroup");
  }
  if (join || !d.phi2L.empty())
    check(ctx, eqb_set_grids(ctx, d.phi2L.data(), d.oma2L.data(), (int)d.phi2L.size(), d.phi2S.data(), d.oma2S.data(),
                             (int)d.phi2S.size()),
          "eqb_set_grids");
  vector<int32_t> gchr(G), schr(M);
  vector<int64_t> gstart(G), gend(G), spos(M), cb(G), ce(G);
  for (int64_t g = 0; g < G; ++g) {
    gchr[g] = chr_idx[genes[g]->chr];
    gstart[g] = genes[g]->start;
    gend[g] = genes[g]->end;
  }
  for (int64_t m = 0; m < M; ++m) {
    schr[m] = chr_idx[snps[m]->chr];
    spos[m] = snps[m]->pos;
  }
  check(ctx, eqb_build_cis_windows(ctx, gchr.data(), gstart.data(), gend.data(), schr.data(), spos.data(),
                                   o.anchor == "TSS" ? EQB_ANCHOR_TSS : EQB_ANCHOR_TSS_TES, (int64_t)o.radius, cb.data(),
                                   ce.data()),
        "eqb_build_cis_windows");
  // genotypes last: the upload is asynchronous (row chunks), projection and results overlap with it
  // Compact lossless transport when the parsed dosages allow it: hard calls (VCF GT, integer dose) travel as u8,
  // dosages written with d <= 4 decimals as u8/u16 numerators of 10^d (k / 10^d correctly rounded = the parsed
  // double, checked element by element below); anything else (NaN, more decimals) goes up as doubles.
  vector<vector<uint8_t> > Gfix8(Gmats.size());
  vector<vector<uint16_t> > Gfix16(Gmats.size());
  for (size_t j = 0; j < Gmats.size(); ++j) {
    const vector<double> &Gm = Gmats[j];
    int width = 0;
    double denom = 1.0;
    for (int dgt = 0; dgt <= 4 && !width; ++dgt, denom *= 10.0) {
      double mx = 0.0;
      bool ok = true;
      for (size_t i = 0; i < Gm.size() && ok; ++i) {
        const double k = nearbyint(Gm[i] * denom);
        ok = (k >= 0.0 && k <= 65535.0 && k / denom == Gm[i]);
        if (k > mx) mx = k;
      }
      if (ok) width = (mx <= 255.0) ? 1 : 2;
      if (width) break;
    }
    if (width == 1) {
      Gfix8[j].resize(Gm.size());
      for (size_t i = 0; i < Gm.size(); ++i) Gfix8[j][i] = (uint8_t)nearbyint(Gm[i] * denom);
      check(ctx, eqb_set_genotypes_fixed(ctx, (int)j, Gfix8[j].data(), 1, denom, M, Gcols[j]), "eqb_set_genotypes_fixed");
    } else if (width == 2) {
      Gfix16[j].resize(Gm.size());
      for (size_t i = 0; i < Gm.size(); ++i) Gfix16[j][i] = (uint16_t)nearbyint(Gm[i] * denom);
      check(ctx, eqb_set_genotypes_fixed(ctx, (int)j, Gfix16[j].data(), 2, denom, M, Gcols[j]), "eqb_set_genotypes_fixed");
    } else
      check(ctx, eqb_set_genotypes(ctx, (int)j, Gmats[j].data(), M, Gcols[j]), "eqb_set_genotypes");
  }
  pc.mark("matrices+upload");
  check(ctx, eqb_finalize(ctx), "eqb_finalize");
  pc.mark("finalize");

  // ---- headers (writeRes(..., "only"), eqtlbma_bf.cpp:1510-1513)
  const int L = (int)d.phi2L.size(), K = (int)d.phi2S.size();
  const int64_t C = eqb_n_configs(ctx);
  const vector<string> cnames = join ? config_names(S, o.bfs) : vector<string>();
  const bool write_ss = !join || (o.outss && o.error != "mvlr");
  const bool is_perm = o.nb_permutations > 0 && (o.perm_sep != 0 || o.pbf != "none");
  const string sep = "\t";
  // (shards after the first one start their files empty: the launcher concatenates the shards' gzip members in order)
  auto gz_header = [&](const string &path, const string &txt) {
    if (o.no_header) {
      FILE *f = fopen(path.c_str(), "wb");
      if (f == NULL) {
        cerr << "ERROR: can't open file " << path << " with mode wb" << endl;
        eqb_exit(EXIT_FAILURE);
      }
      fclose(f);
    } else
      gz_write(path, "wb", txt);
  };
  if (write_ss)
    for (int s = 0; s < S; ++s)
      gz_header(o.out + "_sumstats_" + d.subgroups[s] + ".txt.gz",
                "gene\tsnp\tmaf\tn\tpve\tsigmahat\tbetahat.geno\tsebetahat.geno\tbetapval.geno\n");
  if (join) {
    string h = "gene\tsnp\tconfig";
    for (int i = 0; i < L; ++i) h += "\tl10abf.grid" + to_string(i + 1);
    gz_header(o.out + "_l10abfs_raw.txt.gz", h + "\n");
    if (o.outw) {
      h = "gene\tsnp\tnb.subgroups\tl10abf.gen\tl10abf.gen.fix\tl10abf.gen.maxh";
      if (o.bfs != "gen") h += "\tl10abf.gen.sin";
      if (o.bfs == "all") h += "\tl10abf.all";
      for (size_t c = 0; c < cnames.size(); ++c) h += "\tl10abf." + cnames[c];
      gz_header(o.out + "_l10abfs_avg-grids.txt.gz", h + "\n");
    }
    if (o.nb_permutations > 0) {
      stringstream ss;
      ss << "# perm.bf=" << o.pbf << " seed=" << o.seed << "\n"
         << "gene\tnb.snps\tjoin.perm.pval\tnb.permutations\ttrue.l10abf\tmed.perm.l10abf\n";
      gz_header(o.out + "_joinPermPvals.txt.gz", ss.str());
    }
  } else if (o.nb_permutations > 0 && o.perm_sep != 0) {
    stringstream ss;
    ss << "# seed=" << o.seed << "\n"
       << "gene\tnb.snps\tsep.perm.pval\tnb.permutations\ttrue.min.pval\n";
    if (o.perm_sep == 1)
      gz_header(o.out + "_sepPermPvals.txt.gz", ss.str());
    else
      for (int s = 0; s < S; ++s) gz_header(o.out + "_sepPermPvals_" + d.subgroups[s] + ".txt.gz", ss.str());
  }

  if (o.verbose > 0)
    cout << "test for association between each pair gene-SNP ..." << endl
         << "analysis=" << o.analys << " likelihood=" << o.lik << " error_model=" << o.error << " anchor=" << o.anchor
         << " radius=" << o.radius << endl
         << flush;

  // ---- batches of whole write-groups, sized by a host-memory budget for the raw ABFs
  const size_t per_pair = (size_t)S * 44 + (join ? ((size_t)3 * L + (size_t)C * K + 5 + C) * 8 : 0);
  const size_t budget = (size_t)1 << 30;
  // permutation statistics kept per gene of a batch (device + host copy for the median): [per][nperm] doubles
  const size_t per_gene = o.nb_permutations > 0
                              ? (size_t)((o.analys == "sep" && o.perm_sep == 2) ? S : 1) * (size_t)o.nb_permutations * 16
                              : 0;
  size_t nbAnalyzedGenes = 0, nbAnalyzedPairs = 0;
  int64_t g0 = 0, g_end = G;
  if (o.shard_n > 1) { // gene sharding over GPUs: contiguous cost-balanced ranges of whole write-groups
    vector<int64_t> cost(G), sb(o.shard_n + 1);
    for (int64_t g = 0; g < G; ++g) cost[g] = (ce[g] - cb[g]) * (int64_t)(1 + o.nb_permutations);
    if (eqb_partition_by_cost(cost.data(), G, o.wrtsize, o.shard_n, sb.data()) != 0) {
      cerr << "ERROR: eqb_partition_by_cost failed" << endl;
      eqb_exit(EXIT_FAILURE);
    }
    g0 = sb[o.shard_k];
    g_end = sb[o.shard_k + 1];
  }
  while (g0 < g_end) {
    int64_t g1 = g0;
    size_t pairs_est = 0;
    while (g1 < g_end) {
      int64_t gn = min<int64_t>(g_end, g1 + o.wrtsize);
      size_t add = 0;
      for (int64_t g = g1; g < gn; ++g) add += (size_t)(ce[g] - cb[g]);
      if (g1 > g0 && (pairs_est + add) * per_pair + (size_t)(gn - g0) * per_gene > budget) break;
      pairs_est += add;
      g1 = gn;
    }
    vector<int64_t> off(g1 - g0 + 1);
    check(ctx, eqb_pair_offsets(ctx, g0, g1, off.data()), "eqb_pair_offsets");
    const int64_t P = off[g1 - g0];
    vector<int32_t> n((size_t)P * S);
    vector<double> ss((size_t)P * S * 5), agen, acfg, aw;
    vector<uint8_t> analyzed(g1 - g0);
    eqb_results res;
    memset(&res, 0, sizeof(res));
    res.n = n.data();
    res.sstats = ss.data();
    res.gene_analyzed = analyzed.data();
    if (join) {
      agen.resize((size_t)P * 3 * L);
      acfg.resize((size_t)P * C * K);
      aw.resize((size_t)P * (5 + C));
      res.abf_gen = agen.data();
      res.abf_cfg = acfg.data();
      res.abf_w = aw.data();
    }
    pc.mark("other");
    check(ctx, eqb_run(ctx, g0, g1, &res), "eqb_run");
    pc.mark("run");
    const int per = (!join && o.perm_sep == 2) ? S : 1;
    vector<double> pv, ptrue, pmed;
    vector<int64_t> pdone, pcount;
    if (is_perm) {
      eqb_perm_config pc;
      memset(&pc, 0, sizeof(pc));
      pc.nperm = (int64_t)o.nb_permutations;
      pc.seed = (uint64_t)o.seed;
      pc.trick = o.trick;
      pc.tricut = (int)o.tricut;
      pc.permsep = o.perm_sep;
      pc.pbf = o.pbf == "gen" ? EQB_PBF_GEN : (o.pbf == "gen-sin" ? EQB_PBF_GEN_SIN : (o.pbf == "all" ? EQB_PBF_ALL : EQB_PBF_NONE));
      pc.maxbf = o.maxbf ? 1 : 0;
      pc.wrtsize = o.wrtsize;
      pv.assign((size_t)(g1 - g0) * per, kNaN);
      ptrue.assign((size_t)(g1 - g0) * per, kNaN);
      pmed.assign((size_t)(g1 - g0) * per, kNaN);
      pdone.assign((size_t)(g1 - g0) * per, 0);
      pcount.assign((size_t)(g1 - g0) * per, 0);
      eqb_perm_results pr;
      memset(&pr, 0, sizeof(pr));
      pr.pval = pv.data();
      pr.nperm_done = pdone.data();
      pr.count = pcount.data();
      pr.true_stat = ptrue.data();
      pr.median_perm = pmed.data();
      check(ctx, eqb_run_permutations(ctx, g0, g1, &pc, &pr), "eqb_run_permutations");
      // NaN permuted statistics shrink a gene's number of permutations.  The reference decrements the caller's counter by
      // reference (gene.cpp:431,551,681,693), so that every LATER gene of the run also loses them; here the loss stays with
      // the gene: say so when it happens, the p-values of later genes differ from the reference's then
      if (o.trick == 0) {
        static bool warned = false;
        for (size_t i = 0; i < pdone.size() && !warned; ++i)
          if (pdone[i] > 0 && pdone[i] < (int64_t)o.nb_permutations) {
            cerr << "WARNING: gene " << genes[g0 + (int64_t)(i / per)]->name << ": " << (o.nb_permutations - pdone[i])
                 << " permuted statistics are NaN; they are dropped for this gene only (the reference would also drop as many"
                 << " permutations from every later gene)" << endl;
            warned = true;
          }
      }
    }

    // ---- serialisation (writeRes*, eqtlbma_bf.cpp:919-1399)
    // --thread is repurposed for the output encoder (formatting + deflate), SURVEY 8b
    const int nthr = o.nb_threads;
    if (write_ss) {
      for (int s = 0; s < S; ++s) {
        parallel_emit(o.out + "_sumstats_" + d.subgroups[s] + ".txt.gz", nthr, g0, g1, off,
                      [&](int64_t ga, int64_t gb, string &txt) {
          for (int64_t g = ga; g < gb; ++g) {
            if (!analyzed[g - g0]) continue;
            for (int64_t j = 0; j < ce[g] - cb[g]; ++j) {
              const int64_t p = off[g - g0] + j;
              if (n[p * S + s] <= 0) continue;
              const SnpRec *sr = snps[cb[g] + j];
              txt += genes[g]->name;
              txt += sep;
              txt += sr->name;
              txt += sep;
              put_sci(txt, sr->maf.find(d.subgroups[s])->second);
              txt += sep;
              txt += to_string(n[p * S + s]);
              for (int k = 0; k < 5; ++k) {
                txt += sep;
                put_sci(txt, ss[(p * S + s) * 5 + k]);
              }
              txt += "\n";
            }
          }
        });
      }
    }
    if (join) {
      static const char *rows[3] = {"gen", "gen-fix", "gen-maxh"};
      parallel_emit(o.out + "_l10abfs_raw.txt.gz", nthr, g0, g1, off, [&](int64_t ga, int64_t gb, string &raw) {
        for (int64_t g = ga; g < gb; ++g) {
          if (!analyzed[g - g0]) continue;
          for (int64_t j = 0; j < ce[g] - cb[g]; ++j) {
            const int64_t p = off[g - g0] + j;
            const string &gn = genes[g]->name, &sn = snps[cb[g] + j]->name;
            for (int r = 0; r < 3; ++r) {
              raw += gn + sep + sn + sep + rows[r];
              for (int k = 0; k < L; ++k) {
                raw += sep;
                put_sci(raw, agen[(p * 3 + r) * L + k]);
              }
              raw += "\n";
            }
            for (int64_t c = 0; c < C; ++c) {
              raw += gn + sep + sn + sep + cnames[c];
              for (int k = 0; k < L; ++k) { // padded / truncated to |gridL| columns (eqtlbma_bf.cpp:1207-1212)
                raw += sep;
                put_sci(raw, k < K ? acfg[(p * C + c) * K + k] : kNaN);
              }
              raw += "\n";
            }
          }
        }
      });
      if (o.outw)
        parallel_emit(o.out + "_l10abfs_avg-grids.txt.gz", nthr, g0, g1, off, [&](int64_t ga, int64_t gb, string &avg) {
          for (int64_t g = ga; g < gb; ++g) {
            if (!analyzed[g - g0]) continue;
            for (int64_t j = 0; j < ce[g] - cb[g]; ++j) {
              const int64_t p = off[g - g0] + j;
              int nsub = 0;
              for (int s = 0; s < S; ++s) nsub += n[p * S + s] > 0 ? 1 : 0;
              avg += genes[g]->name + sep + snps[cb[g] + j]->name + sep + to_string(nsub);
              const double *w = &aw[p * (5 + C)];
              for (int k = 0; k < 3; ++k) {
                avg += sep;
                put_sci(avg, w[k]);
              }
              if (o.bfs != "gen") {
                avg += sep;
                put_sci(avg, w[3]);
              }
              if (o.bfs == "all") {
                avg += sep;
                put_sci(avg, w[4]);
              }
              for (int64_t c = 0; c < C; ++c) {
                avg += sep;
                put_sci(avg, w[5 + c]);
              }
              avg += "\n";
            }
          }
        });
    }
    if (o.nb_permutations > 0 && (join || o.perm_sep != 0)) {
      for (int s = 0; s < per; ++s) {
        string txt;
        for (int64_t g = g0; g < g1; ++g) {
          if (!analyzed[g - g0]) continue;
          const int64_t npairs = ce[g] - cb[g];
          if (npairs <= 0) continue;
          const size_t r = (size_t)(g - g0) * per + s;
          int64_t nsnps = npairs;
          if (per > 1) { // GetNbGeneSnpPairs(subgroup): pairs with results in that subgroup
            nsnps = 0;
            for (int64_t j = 0; j < npairs; ++j) nsnps += n[(off[g - g0] + j) * S + s] > 0 ? 1 : 0;
          }
          txt += genes[g]->name + sep + to_string(nsnps) + sep;
          put_def(txt, is_perm ? pv[r] : kNaN);
          txt += sep + to_string(is_perm ? pdone[r] : 0) + sep;
          put_sci(txt, is_perm ? ptrue[r] : kNaN);
          if (join) {
            txt += sep;
            put_sci(txt, is_perm ? pmed[r] : kNaN);
          }
          txt += "\n";
        }
        const string path = join ? o.out + "_joinPermPvals.txt.gz"
                                 : (per > 1 ? o.out + "_sepPermPvals_" + d.subgroups[s] + ".txt.gz" : o.out + "_sepPermPvals.txt.gz");
        gz_write(path, "ab", txt);
      }
    }
    for (int64_t g = g0; g < g1; ++g)
      if (analyzed[g - g0]) {
        ++nbAnalyzedGenes;
        nbAnalyzedPairs += (size_t)(ce[g] - cb[g]);
      }
    g0 = g1;
  }
  if (o.shard_child) { // the launcher adds the shards' counts up
    FILE *f = fopen((o.out + ".count").c_str(), "w");
    if (f != NULL) {
      fprintf(f, "%zu %zu\n", nbAnalyzedPairs, nbAnalyzedGenes);
      fclose(f);
    }
  } else if (o.verbose > 0)
    cout << "nb of analyzed gene-SNP pairs: " << nbAnalyzedPairs << " (" << nbAnalyzedGenes << " genes)" << endl;
  pc.mark("write");
  eqb_destroy(ctx);
  pc.mark("destroy");
  if (o.verbose > 1) pc.print();
  if (o.verbose > 0 && !o.shard_child) {
    time_t t_end;
    time(&t_end);
    cout << "END " << argv[0] << " " << ctime(&t_end) << "elapsed -> " << difftime(t_end, t_start) << " sec" << endl;
  }
  return EXIT_SUCCESS;
}
