// eqtlbma_hm -- drop-in front-end of the reference's `eqtlbma_hm` (src/eqtlbma_hm.cpp:1756-2151) for `--model configs`,
// above the C ABI of include/eqtlbma_hm_b200.h: same options, same input (`_l10abfs_raw.txt.gz` files of eqtlbma_bf), same
// output file (save_result, eqtlbma_hm.cpp:1612-1752).  Every number is computed by libeqtlbma_b200.so on the GPU; the host
// does what the reference's host does around the EM: parse, initialise, print, write.
//
// Loader: each file is inflated once, its lines are indexed serially (gene / snp / config tokens, the `--configs` /
// `--keepgen` filters of load_data_one_file, eqtlbma_hm.cpp:315-332) and the numeric cells are parsed by all host threads
// straight into the [pair][config][grid point] array that eqb_hm_append takes.
// Not covered (rejected with a message): --model types.
#include <getopt.h>
#include <glob.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../../include/eqtlbma_b200.h"
#include "../../include/eqtlbma_hm_b200.h"

namespace {

const double NaN = std::numeric_limits<double>::quiet_NaN();

struct Options {
  int verbose = 1, threads = 1, gpus = 1;
  std::string data, model = "configs", out, init, ci;
  long nsubgrp = -1, dim = -1, ngrid = -1, maxit = -1;
  bool rand_init = false, has_seed = false, keepgen = false, getci = false, getbf = false;
  unsigned long seed = 0;
  double thresh = 0.05, msl = 1.0, pi0 = NaN;
  std::vector<std::string> configs;
};

std::thread *g_warm = nullptr; // CUDA context creation in flight: every exit path waits for it
void join_warmup()
{
  if (g_warm && g_warm->joinable()) g_warm->join();
}

[[noreturn]] void die(const std::string &msg)
{
  fprintf(stderr, "ERROR: %s\n", msg.c_str());
  join_warmup();
  exit(EXIT_FAILURE);
}

void help(const char *argv0)
{
  printf("`%s' fits the hierarchical model of eQtlBma with an EM algorithm (B200 build, --model configs).\n\n"
         "Usage: %s [OPTIONS] ...\n\nOptions:\n"
         "  -h, --help\tdisplay the help and exit\n  -V, --version\toutput version information and exit\n"
         "  -v, --verbose\tverbosity level (0/default=1/2/3)\n"
         "      --data\tinput data (usually output files from eqtlbma_bf)\n      --nsubgrp\tnumber of subgroups\n"
         "      --model\twhich model to fit (only configs here)\n      --dim\tdimension of the model (nb of active configs)\n"
         "      --ngrid\tnumber of grid points\n      --out\toutput file (gzipped)\n      --init\tfile for initialization\n"
         "\t\t3 columns: param<tab>value<tab>fixed (TRUE or FALSE)\n      --rand\trandom initialization\n"
         "      --seed\tseed used with --rand, otherwise use time\n      --thresh\tthreshold to stop the EM (default=0.05)\n"
         "      --maxit\tmaximum number of iterations (optional)\n      --msl\tmaximum step length for SQUAREM (default=1: classical EM)\n"
         "      --thread\tnumber of host threads of the loader (default=1; the EM itself runs on the GPU)\n"
         "      --configs\tsubset of configurations to keep (e.g. \"1|3|1-3\")\n      --keepgen\tkeep 'general' ABFs\n"
         "      --getci\tcompute the confidence intervals\n      --getbf\tcompute the Bayes Factors using the estimated weights\n"
         "      --pi0\tfixed value for pi0\n      --ci\tfile with estimates of hyperparameters to only compute confidence intervals\n"
         "      --gpus\tnumber of GPUs (default=1): one process per GPU, the input files (or the genes) cut into contiguous shards,\n"
         "\t\tthe sums over genes exchanged over NVLink peer memory; same output file\n",
         argv0, argv0);
}

Options parse_cmdline(int argc, char **argv)
{
  Options o;
  static struct option lo[] = {{"help", no_argument, 0, 'h'},       {"version", no_argument, 0, 'V'},  {"verbose", required_argument, 0, 'v'},
                               {"data", required_argument, 0, 0},   {"nsubgrp", required_argument, 0, 0}, {"model", required_argument, 0, 0},
                               {"dim", required_argument, 0, 0},    {"ngrid", required_argument, 0, 0}, {"out", required_argument, 0, 0},
                               {"init", required_argument, 0, 0},   {"rand", no_argument, 0, 0},        {"seed", required_argument, 0, 0},
                               {"thresh", required_argument, 0, 0}, {"maxit", required_argument, 0, 0}, {"msl", required_argument, 0, 0},
                               {"thread", required_argument, 0, 0}, {"configs", required_argument, 0, 0}, {"keepgen", no_argument, 0, 0},
                               {"getci", no_argument, 0, 0},        {"getbf", no_argument, 0, 0},       {"pi0", required_argument, 0, 0},
                               {"ci", required_argument, 0, 0},     {"gpus", required_argument, 0, 0},  {0, 0, 0, 0}};
  int c, idx = 0;
  while ((c = getopt_long(argc, argv, "hVv:", lo, &idx)) != -1) {
    if (c == 'h') {
      help(argv[0]);
      exit(0);
    } else if (c == 'V') {
      printf("%s (eqtlbma_b200)\n", argv[0]);
      exit(0);
    } else if (c == 'v')
      o.verbose = atoi(optarg);
    else if (c == 0) {
      const std::string n = lo[idx].name;
      if (n == "data") o.data = optarg;
      else if (n == "nsubgrp") o.nsubgrp = atol(optarg);
      else if (n == "model") o.model = optarg;
      else if (n == "dim") o.dim = atol(optarg);
      else if (n == "ngrid") o.ngrid = atol(optarg);
      else if (n == "out") o.out = optarg;
      else if (n == "init") o.init = optarg;
      else if (n == "rand") o.rand_init = true;
      else if (n == "seed") {
        o.seed = (unsigned long)atol(optarg);
        o.has_seed = true;
      } else if (n == "thresh") o.thresh = atof(optarg);
      else if (n == "maxit") o.maxit = atol(optarg);
      else if (n == "msl") o.msl = atof(optarg);
      else if (n == "thread") o.threads = std::max(1, atoi(optarg));
      else if (n == "configs") {
        std::string s = optarg;
        size_t p = 0;
        while (true) {
          const size_t q = s.find('|', p);
          const std::string t = s.substr(p, q == std::string::npos ? q : q - p);
          if (!t.empty()) o.configs.push_back(t);
          if (q == std::string::npos) break;
          p = q + 1;
        }
      } else if (n == "keepgen") o.keepgen = true;
      else if (n == "getci") o.getci = true;
      else if (n == "getbf") o.getbf = true;
      else if (n == "pi0") o.pi0 = atof(optarg);
      else if (n == "ci") o.ci = optarg;
      else if (n == "gpus") o.gpus = std::max(1, atoi(optarg));
    } else
      exit(EXIT_FAILURE);
  }
  // the reference's checks (eqtlbma_hm.cpp:1957-2050)
  if (o.data.empty()) die("missing compulsory option --data");
  if (o.nsubgrp <= 0) die("missing compulsory option --nsubgrp");
  if (o.model != "configs") die("--model " + o.model + " is not covered by the B200 build (only configs)");
  if (o.dim <= 0) die("missing compulsory option --dim");
  if (o.ngrid <= 0) die("missing compulsory option --ngrid");
  if (o.out.empty()) die("missing compulsory option --out");
  if (!o.init.empty() && o.rand_init) die("--init and --rand are mutually exclusive");
  if (o.rand_init && !o.has_seed) {
    o.seed = (unsigned long)std::chrono::system_clock::now().time_since_epoch().count() % 1000000007ul;
    o.has_seed = true;
  }
  if (!std::isnan(o.pi0) && (o.pi0 <= 0.0 || o.pi0 >= 1.0)) die("--pi0 is invalid");
  if (!o.ci.empty() && !o.getci) die("--ci should be used with --getci");
  return o;
}

// ---------------------------------------------------------------------------------------------------------------- loader
struct Data {
  std::vector<std::string> gene_names, snp_names, config_names;
  std::vector<int64_t> gene_off{0};
  std::vector<double> B; // [pairs][dim][grid]
};

void inflate_file(const std::string &path, std::string &buf)
{
  gzFile f = gzopen(path.c_str(), "rb");
  if (!f) die("can't open file " + path);
  gzbuffer(f, 1 << 20);
  buf.clear();
  std::vector<char> chunk(1 << 22);
  int n;
  while ((n = gzread(f, chunk.data(), (unsigned)chunk.size())) > 0) buf.append(chunk.data(), (size_t)n);
  if (n < 0) die("can't read file " + path);
  gzclose(f);
}

inline bool is_sep(char c) { return c == ' ' || c == '\t' || c == ','; }
inline const char *skip_seps(const char *p, const char *e)
{
  while (p < e && is_sep(*p)) ++p;
  return p;
}
inline const char *token_end(const char *p, const char *e)
{
  while (p < e && !is_sep(*p)) ++p;
  return p;
}

// one file (load_data_one_file, eqtlbma_hm.cpp:287-371): rows are kept in file order; a new gene starts when the gene token
// changes, a new SNP when the SNP token changes
void load_one_file(const std::string &path, const Options &o, Data &d, std::string &cur_gene, std::string &cur_snp, int &cfg_in_pair)
{
  std::string buf;
  inflate_file(path, buf);
  const char *p = buf.data(), *end = p + buf.size();
  const char *eol = (const char *)memchr(p, '\n', (size_t)(end - p));
  if (!eol) eol = end;
  {
    const char *a = skip_seps(p, eol), *ae = token_end(a, eol);
    const char *b = skip_seps(ae, eol), *be = token_end(b, eol);
    const char *c = skip_seps(be, eol), *ce = token_end(c, eol);
    if (std::string(a, ae) != "gene" || std::string(b, be) != "snp" || std::string(c, ce) != "config") die("file " + path + " has wrong header line");
  }
  struct Row {
    const char *cells, *eol;
    size_t row; // index of the (pair, config) row in B
  };
  std::vector<Row> rows;
  const size_t dim = (size_t)o.dim, grid = (size_t)o.ngrid;
  size_t n_rows = d.B.size() / grid;
  bool first_in_file = true; // the reference's current gene / SNP are locals of load_data_one_file: a file starts a gene
  p = (eol < end) ? eol + 1 : end;
  while (p < end) {
    eol = (const char *)memchr(p, '\n', (size_t)(end - p));
    if (!eol) eol = end;
    const char *a = skip_seps(p, eol), *ae = token_end(a, eol);
    if (a == ae) { // empty line
      p = eol + 1;
      continue;
    }
    const char *b = skip_seps(ae, eol), *be = token_end(b, eol);
    const char *c = skip_seps(be, eol), *ce = token_end(c, eol);
    const std::string config(c, ce);
    const bool has_gen = config.find("gen") != std::string::npos;
    bool skip = (!o.keepgen && has_gen) ||
                (!o.configs.empty() && std::find(o.configs.begin(), o.configs.end(), config) == o.configs.end()) ||
                (o.keepgen && (config == "gen-fix" || config == "gen-maxh"));
    if (!skip) {
      if (d.config_names.size() < dim) d.config_names.push_back(config);
      const bool new_gene = cur_gene.size() != (size_t)(ae - a) || memcmp(cur_gene.data(), a, (size_t)(ae - a)) != 0 || first_in_file;
      first_in_file = false;
      const bool new_snp = new_gene || cur_snp.size() != (size_t)(be - b) || memcmp(cur_snp.data(), b, (size_t)(be - b)) != 0;
      if (new_gene) {
        cur_gene.assign(a, ae);
        d.gene_names.push_back(cur_gene);
        d.gene_off.push_back(d.gene_off.back());
      }
      if (new_snp) {
        if (!d.snp_names.empty() && cfg_in_pair != (int)dim)
          die("snp " + d.snp_names.back() + " has " + std::to_string(cfg_in_pair) + " configurations instead of --dim " + std::to_string(dim));
        cur_snp.assign(b, be);
        d.snp_names.push_back(cur_snp);
        ++d.gene_off.back();
        cfg_in_pair = 0;
      }
      if (cfg_in_pair >= (int)dim) die("snp " + cur_snp + " has more configurations than --dim " + std::to_string(dim));
      ++cfg_in_pair;
      rows.push_back(Row{ce, eol, n_rows++});
    }
    p = eol + 1;
  }
  d.B.resize(n_rows * grid, NaN);
  // numeric cells, in parallel (atof semantics: strtod of each token; missing cells stay NaN as in the reference)
  const int nt = std::max(1, std::min<int>(o.threads, (int)(rows.size() / 4096) + 1));
  auto work = [&](int t) {
    const size_t lo = rows.size() * (size_t)t / (size_t)nt, hi = rows.size() * (size_t)(t + 1) / (size_t)nt;
    for (size_t i = lo; i < hi; ++i) {
      const char *q = rows[i].cells, *e = rows[i].eol;
      double *out = d.B.data() + rows[i].row * grid;
      for (size_t j = 0; j < grid; ++j) {
        q = skip_seps(q, e);
        if (q >= e) break;
        char *stop = nullptr;
        out[j] = strtod(q, &stop); // the buffer ends with the file's last newline or a NUL of std::string: never overruns
        q = token_end(q, e);
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
  work(0);
  for (auto &x : th) x.join();
}

// ------------------------------------------------------------------------------------------------ initial values, RNG
// gsl_rng_mt19937 as seeded by gsl_rng_set (GSL rng/mt.c), gsl_rng_uniform = get / 2^32, gsl_ran_exponential = -mu log1p(-u)
struct Mt19937 {
  uint32_t mt[624];
  int mti;
  explicit Mt19937(unsigned long s)
  {
    if (s == 0) s = 4357;
    mt[0] = (uint32_t)(s & 0xffffffffUL);
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    mti = 624;
  }
  uint32_t get()
  {
    if (mti >= 624) {
      int kk;
      for (kk = 0; kk < 624 - 397; ++kk) {
        const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      for (; kk < 623; ++kk) {
        const uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      const uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
      mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      mti = 0;
    }
    uint32_t k = mt[mti++];
    k ^= (k >> 11);
    k ^= (k << 7) & 0x9d2c5680u;
    k ^= (k << 15) & 0xefc60000u;
    k ^= (k >> 18);
    return k;
  }
  double uniform() { return get() / 4294967296.0; }
  double exponential(double mu) { return -mu * log1p(-uniform()); }
};

struct Params {
  double pi0 = NaN;
  std::vector<double> grid, config;
  bool fixed_pi0 = false, fixed_grid = false, fixed_configs = false;
};

// Controller::init_params(file) (eqtlbma_hm.cpp:452-527)
void init_from_file(const std::string &path, Params &pr, int verbose)
{
  if (verbose) printf("loading initialization file %s ...\n", path.c_str());
  std::string buf;
  inflate_file(path, buf); // (gzopen reads plain text too)
  size_t ig = 0, ic = 0, p = 0;
  while (p < buf.size()) {
    size_t e = buf.find('\n', p);
    if (e == std::string::npos) e = buf.size();
    const std::string line = buf.substr(p, e - p);
    p = e + 1;
    if (line.empty() || line[0] == '#') {
      // the reference's own output file carries the estimates on `#`-prefixed lines: test_hm.bash strips the `#` first
      continue;
    }
    std::vector<std::string> tok;
    size_t a = 0;
    while (a < line.size()) {
      while (a < line.size() && is_sep(line[a])) ++a;
      size_t b = a;
      while (b < line.size() && !is_sep(line[b])) ++b;
      if (b > a) tok.push_back(line.substr(a, b - a));
      a = b;
    }
    if (tok.size() < 2 || (tok[0] == "param" && tok[1] == "value")) continue;
    // (the reference tests tokens.size() == 3: files with the 5 columns of its own output mark nothing as fixed)
    const bool fx = tok.size() == 3 && (tok[2] == "TRUE" || tok[2] == "true");
    if (tok[0].find("pi0") != std::string::npos) {
      pr.pi0 = atof(tok[1].c_str());
      if (fx) pr.fixed_pi0 = true;
    } else if (tok[0].find("grid") != std::string::npos) {
      if (ig < pr.grid.size()) pr.grid[ig] = atof(tok[1].c_str());
      ++ig;
      if (fx) pr.fixed_grid = true;
    } else if (tok[0].find("config") != std::string::npos) {
      if (ic < pr.config.size()) pr.config[ic] = atof(tok[1].c_str());
      ++ic;
      if (fx) pr.fixed_configs = true;
    }
  }
  if (verbose > 0) {
    if (!pr.fixed_pi0 && !pr.fixed_grid && !pr.fixed_configs)
      printf("update all parameters\n");
    else {
      printf("parameters to update:");
      if (!pr.fixed_configs) printf(" configs");
      if (!pr.fixed_grid) printf(" grid-points");
      if (!pr.fixed_pi0) printf(" pi0");
      printf("\n");
    }
  }
}

// Controller::init_params(seed) (eqtlbma_hm.cpp:529-613)
void init_default(const Options &o, Params &pr)
{
  const size_t dim = pr.config.size(), grid = pr.grid.size();
  if (o.has_seed) {
    Mt19937 r(o.seed);
    if (!pr.fixed_pi0) pr.pi0 = r.uniform();
    double sum = 0;
    for (size_t l = 0; l < grid; ++l) sum += (pr.grid[l] = r.exponential(1.0));
    for (size_t l = 0; l < grid; ++l) pr.grid[l] /= sum;
    if (dim == 1)
      pr.config[0] = 1.0;
    else {
      sum = 0;
      for (size_t k = 0; k < dim; ++k) sum += (pr.config[k] = r.exponential(1.0));
      for (size_t k = 0; k < dim; ++k) pr.config[k] /= sum;
    }
  } else {
    if (!pr.fixed_pi0) pr.pi0 = 0.5;
    for (size_t l = 0; l < grid; ++l) pr.grid[l] = 1.0 / (double)grid;
    for (size_t k = 0; k < dim; ++k) pr.config[k] = 1.0 / (double)dim;
  }
}

// ---------------------------------------------------------------------------------------------------------------- writer
struct GzOut {
  gzFile f;
  std::string buf;
  explicit GzOut(const std::string &path) : f(gzopen(path.c_str(), "wb"))
  {
    if (!f) die("can't open file " + path);
    gzbuffer(f, 1 << 20);
  }
  void flush()
  {
    if (!buf.empty() && gzwrite(f, buf.data(), (unsigned)buf.size()) <= 0) die("can't write the output file");
    buf.clear();
  }
  void put(const char *s) { buf += s; }
  void num(double v) // operator<< with scientific, precision 4
  {
    char b[48];
    if (std::isnan(v))
      snprintf(b, sizeof(b), "%s", std::signbit(v) ? "-nan" : "nan");
    else if (std::isinf(v))
      snprintf(b, sizeof(b), "%s", v < 0 ? "-inf" : "inf");
    else
      snprintf(b, sizeof(b), "%.4e", v);
    buf += b;
    if (buf.size() > (1u << 22)) flush();
  }
  ~GzOut()
  {
    flush();
    gzclose(f);
  }
};

void hm_log(void *, const char *text)
{
  fputs(text, stdout);
  fflush(stdout);
}

double seconds_since(const std::chrono::steady_clock::time_point &t0)
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

} // namespace

// one process = one GPU: rank `rank` of `world`; xdir = directory through which the ranks exchange their IPC handles
int run_one(Options o, int rank, int world, const std::string &xdir)
{
  const auto t_start = std::chrono::steady_clock::now();
  if (rank > 0) o.verbose = 0; // rank 0 speaks for all (the ranks hold identical estimates)
  // the CUDA context (0.5-1 s on a cold process) is created while the files are parsed
  std::thread warm([rank] { eqb_warmup(rank); });
  g_warm = &warm;

  // ---- load_data (eqtlbma_hm.cpp:373-450)
  if (o.verbose > 0) fprintf(stderr, "load data ...\n");
  if (!o.configs.empty()) {
    fprintf(stderr, "configurations to keep: %s", o.configs[0].c_str());
    for (size_t i = 1; i < o.configs.size(); ++i) fprintf(stderr, " %s", o.configs[i].c_str());
    fprintf(stderr, "\n");
  }
  glob_t gl;
  memset(&gl, 0, sizeof(gl));
  if (glob(o.data.c_str(), 0, nullptr, &gl) != 0 || gl.gl_pathc == 0) die("no input file was found from pattern " + o.data);
  if (o.verbose > 0) printf("nb of input files: %zu\n", (size_t)gl.gl_pathc);
  Data d;
  std::string cur_gene, cur_snp;
  int cfg_in_pair = 0;
  // several GPUs: contiguous shards of the input files when there are enough of them (a file starts a gene, so any cut
  // between files is a cut between genes), of the genes otherwise
  const size_t F = gl.gl_pathc;
  const bool by_file = world > 1 && F >= (size_t)world;
  const size_t f0 = by_file ? F * (size_t)rank / (size_t)world : 0, f1 = by_file ? F * (size_t)(rank + 1) / (size_t)world : F;
  for (size_t i = f0; i < f1; ++i) {
    if (o.verbose > 1) printf("file %zu %s\n", i + 1, gl.gl_pathv[i]);
    load_one_file(gl.gl_pathv[i], o, d, cur_gene, cur_snp, cfg_in_pair);
  }
  globfree(&gl);
  if (d.snp_names.empty()) die("no gene-snp pair was loaded");
  if (cfg_in_pair != (int)o.dim) die("snp " + d.snp_names.back() + " has " + std::to_string(cfg_in_pair) + " configurations instead of --dim " + std::to_string(o.dim));
  if (world > 1 && !by_file) { // keep the genes of this rank
    const int64_t Gall = (int64_t)d.gene_names.size();
    if (Gall < world) die("--gpus " + std::to_string(world) + " needs at least as many genes");
    const int64_t ga = Gall * rank / world, gb = Gall * (rank + 1) / world;
    const int64_t pa = d.gene_off[(size_t)ga], pb = d.gene_off[(size_t)gb];
    const size_t per_pair = (size_t)o.dim * (size_t)o.ngrid;
    Data k;
    k.config_names = d.config_names;
    k.gene_names.assign(d.gene_names.begin() + ga, d.gene_names.begin() + gb);
    k.snp_names.assign(d.snp_names.begin() + pa, d.snp_names.begin() + pb);
    k.gene_off.clear();
    for (int64_t g = ga; g <= gb; ++g) k.gene_off.push_back(d.gene_off[(size_t)g] - pa);
    k.B.assign(d.B.begin() + (size_t)pa * per_pair, d.B.begin() + (size_t)pb * per_pair);
    d = std::move(k);
  }
  const int64_t G = (int64_t)d.gene_names.size(), P = (int64_t)d.snp_names.size();
  join_warmup();
  eqb_hm_ctx *hm = nullptr;
  if (eqb_hm_create(&hm, rank, (int32_t)o.dim, (int32_t)o.ngrid) != 0) die(hm ? eqb_hm_last_error(hm) : "eqb_hm_create failed (no CUDA device? there is no CPU fallback)");
  auto ck = [&](int rc) {
    if (rc != 0) {
      fprintf(stderr, "%s\n", eqb_hm_last_error(hm));
      join_warmup();
      exit(EXIT_FAILURE);
    }
  };
  ck(eqb_hm_append(hm, d.B.data(), P, d.gene_off.data(), G));
  ck(eqb_hm_finalize(hm));
  if (world > 1) { // exchange the IPC handles through files, then connect (include/eqtlbma_hm_b200.h)
    char mine[64];
    ck(eqb_hm_ipc_export(hm, mine));
    const std::string tmp = xdir + "/t." + std::to_string(rank), fin = xdir + "/h." + std::to_string(rank);
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f || fwrite(mine, 1, 64, f) != 64) die("can't write " + tmp);
    fclose(f);
    if (rename(tmp.c_str(), fin.c_str()) != 0) die("can't rename " + tmp);
    std::vector<char> all((size_t)world * 64);
    for (int r = 0; r < world; ++r) {
      const std::string path = xdir + "/h." + std::to_string(r);
      const auto t0 = std::chrono::steady_clock::now();
      FILE *g = nullptr;
      while (!(g = fopen(path.c_str(), "rb"))) {
        if (seconds_since(t0) > 600.0) die("rank " + std::to_string(r) + " did not publish its handle");
        usleep(2000);
      }
      if (fread(all.data() + (size_t)r * 64, 1, 64, g) != 64) die("short read of " + path);
      fclose(g);
    }
    ck(eqb_hm_ipc_connect(hm, world, rank, all.data()));
  }
  if (o.verbose > 0) fprintf(stderr, "finish loading %lld genes and %lld gene-snp pairs (%f sec)\n", (long long)G, (long long)P, seconds_since(t_start));

  Params pr;
  pr.grid.assign((size_t)o.ngrid, NaN);
  pr.config.assign((size_t)o.dim, NaN);
  if (!std::isnan(o.pi0)) {
    pr.pi0 = o.pi0;
    pr.fixed_pi0 = true;
  }
  std::vector<double> grid_ci((size_t)o.ngrid * 2, NaN), config_ci((size_t)o.dim * 2, NaN);
  eqb_hm_fit fit;
  memset(&fit, 0, sizeof(fit));
  fit.pi0_ci[0] = fit.pi0_ci[1] = NaN;
  fit.grid_ci = grid_ci.data();
  fit.config_ci = config_ci.data();
  std::vector<double> gene_post, gene_bf, snp_bf, cfg_bf;
  bool with_bf = false;

  auto bind = [&]() {
    fit.pi0 = pr.pi0;
    fit.grid_wts = pr.grid.data();
    fit.config_prior = pr.config.data();
  };
  if (!o.ci.empty()) {
    // only the intervals, around the estimates of the file (run(), eqtlbma_hm.cpp:2080-2086)
    init_from_file(o.ci, pr, o.verbose);
    bind();
    if (o.verbose > 0) printf("compute profile-likelihood confidence intervals ...\n");
    ck(eqb_hm_profile_ci(hm, &fit));
  } else {
    if (!o.init.empty())
      init_from_file(o.init, pr, o.verbose);
    else
      init_default(o, pr);
    bind();
    if (o.verbose > 0) printf("run EM algorithm (%s) ...\n", o.msl > 1.0 ? "square" : "classic");
    eqb_hm_options opt;
    memset(&opt, 0, sizeof(opt));
    opt.thresh = o.thresh;
    opt.maxit = o.maxit;
    opt.stepmax = o.msl;
    opt.fixed_pi0 = pr.fixed_pi0;
    opt.fixed_grid = pr.fixed_grid;
    opt.fixed_configs = pr.fixed_configs;
    opt.verbose = o.verbose;
    opt.log = rank == 0 ? hm_log : nullptr;
    const auto t_em = std::chrono::steady_clock::now();
    ck(eqb_hm_em(hm, &opt, &fit));
    pr.pi0 = fit.pi0;
    if (rank == 0) printf("EM ran for %.3f sec\n", seconds_since(t_em));
    if (o.getbf) {
      if (o.verbose > 0) printf("compute posteriors ...\n");
      gene_post.assign((size_t)G, NaN);
      gene_bf.assign((size_t)G, NaN);
      snp_bf.assign((size_t)P, NaN);
      cfg_bf.assign((size_t)P * (size_t)o.dim, NaN);
      ck(eqb_hm_posteriors(hm, &fit, gene_post.data(), gene_bf.data(), snp_bf.data(), nullptr, cfg_bf.data(), nullptr));
      with_bf = true;
    }
    if (o.getci) {
      if (o.verbose > 0) printf("compute profile-likelihood confidence intervals ...\n");
      ck(eqb_hm_profile_ci(hm, &fit));
    }
  }

  // ---- save_result (eqtlbma_hm.cpp:1612-1752)
  if (o.verbose > 0) printf("save the results in %s ...\n", o.out.c_str());
  const std::string out_path = world > 1 ? o.out + ".shard" + std::to_string(rank) : o.out;
  {
    GzOut w(out_path);
    auto line = [&](const std::string &name, double mle, double l, double r, bool fixed) {
      w.put(name.c_str());
      w.put("\t");
      w.num(mle);
      w.put("\t");
      w.num(l);
      w.put("\t");
      w.num(r);
      w.put(fixed ? "\ttrue\n" : "\tfalse\n");
    };
    if (rank == 0) {
      w.put("#param\tmle\tleft.ci\tright.ci\tfixed\n");
      line("#pi0", fit.pi0, fit.pi0_ci[0], fit.pi0_ci[1], pr.fixed_pi0);
      for (long k = 0; k < o.dim; ++k)
        line("#config." + (k < (long)d.config_names.size() ? d.config_names[(size_t)k] : std::string("?")), pr.config[(size_t)k],
             config_ci[2 * (size_t)k], config_ci[2 * (size_t)k + 1], pr.fixed_configs);
      for (long l = 0; l < o.ngrid; ++l)
        line("#grid." + std::to_string(l + 1), pr.grid[(size_t)l], grid_ci[2 * (size_t)l], grid_ci[2 * (size_t)l + 1], pr.fixed_grid);
    }
    if (with_bf) {
      if (rank == 0) {
        w.put("gene\tgene.posterior.prob\tgene.log10.bf\tsnp\tsnp.log10.bf");
        for (long k = 0; k < o.dim; ++k) {
          w.put("\tlog10.bf.");
          w.put(d.config_names[(size_t)k].c_str());
        }
        w.put("\n");
      }
      for (int64_t g = 0; g < G; ++g)
        for (int64_t p = d.gene_off[(size_t)g]; p < d.gene_off[(size_t)g + 1]; ++p) {
          w.put(d.gene_names[(size_t)g].c_str());
          w.put("\t");
          w.num(gene_post[(size_t)g]);
          w.put("\t");
          w.num(gene_bf[(size_t)g]);
          w.put("\t");
          w.put(d.snp_names[(size_t)p].c_str());
          w.put("\t");
          w.num(snp_bf[(size_t)p]);
          for (long k = 0; k < o.dim; ++k) {
            w.put("\t");
            w.num(cfg_bf[(size_t)p * (size_t)o.dim + (size_t)k]);
          }
          w.put("\n");
        }
    }
  }
  eqb_hm_destroy(hm);
  if (o.verbose > 0) printf("END (%.3f sec)\n", seconds_since(t_start));
  return EXIT_SUCCESS;
}

int main(int argc, char **argv)
{
  const Options o = parse_cmdline(argc, argv);
  if (o.gpus <= 1) return run_one(o, 0, 1, "");
  // --gpus N: one child per GPU, forked before any CUDA call; the children write <out>.shard<k> (gzip members; rank 0 alone
  // writes the parameter lines and the header), concatenated byte-wise in rank order = gene order into <out>
  char tmpl[] = "/tmp/eqtlbma_hm_XXXXXX";
  if (!mkdtemp(tmpl)) die("can't create a temporary directory");
  const std::string xdir = tmpl;
  std::vector<pid_t> kids;
  for (int r = 0; r < o.gpus; ++r) {
    fflush(stdout);
    fflush(stderr);
    const pid_t pid = fork();
    if (pid < 0) die("fork failed");
    if (pid == 0) _exit(run_one(o, r, o.gpus, xdir));
    kids.push_back(pid);
  }
  bool ok = true;
  for (pid_t pid : kids) {
    int st = 0;
    if (waitpid(pid, &st, 0) < 0 || !WIFEXITED(st) || WEXITSTATUS(st) != 0) ok = false;
  }
  for (int r = 0; r < o.gpus; ++r) {
    unlink((xdir + "/h." + std::to_string(r)).c_str());
    unlink((xdir + "/t." + std::to_string(r)).c_str());
  }
  rmdir(xdir.c_str());
  if (ok) {
    FILE *dst = fopen(o.out.c_str(), "wb");
    if (!dst) die("can't open file " + o.out);
    std::vector<char> buf(1 << 20);
    for (int r = 0; r < o.gpus; ++r) {
      const std::string path = o.out + ".shard" + std::to_string(r);
      FILE *src = fopen(path.c_str(), "rb");
      if (!src) die("missing shard output " + path);
      size_t n;
      while ((n = fread(buf.data(), 1, buf.size(), src)) > 0)
        if (fwrite(buf.data(), 1, n, dst) != n) die("can't write " + o.out);
      fclose(src);
    }
    fclose(dst);
  }
  for (int r = 0; r < o.gpus; ++r) unlink((o.out + ".shard" + std::to_string(r)).c_str());
  if (!ok) die("a shard process failed");
  return EXIT_SUCCESS;
}
