/*
 * ref_hm_dump_main.cpp -- TEST INFRASTRUCTURE (oracle side), never part of the product path.
 *
 * Drives the UNMODIFIED reference eqtlbma_hm (it textually includes the reference's own
 * src/eqtlbma_hm.cpp from where it lies under /root/reference, with `main` renamed) through the
 * same sequence as its run() (eqtlbma_hm.cpp:2053-2104) and writes, in addition to the
 * reference's own output file, every fitted quantity at FULL precision (%.17g) into the file
 * named by $EQTLBMA_HM_DUMP.  The reference prints 4 significant digits only
 * (eqtlbma_hm.cpp:1632); the dump is what the parity checks of the EM path are pinned against.
 */
#define main eqtlbma_hm_reference_main
#include REF_HM_CPP
#undef main

static void dumpVec(FILE *f, const char *tag, const vector<double> &v)
{
  fprintf(f, "%s", tag);
  for (size_t i = 0; i < v.size(); ++i) fprintf(f, "\t%.17g", v[i]);
  fprintf(f, "\n");
}

static void dumpAll(Controller &c, bool with_bf)
{
  const char *path = getenv("EQTLBMA_HM_DUMP");
  if (!path) return;
  FILE *f = fopen(path, "w");
  if (!f) {
    perror(path);
    exit(EXIT_FAILURE);
  }
  fprintf(f, "SHAPE\t%zu\t%zu\t%zu\n", c.genes_.size(), c.dim_, c.grid_size_);
  fprintf(f, "LOGLIK\t%.17g\n", c.log10_obs_lik_);
  fprintf(f, "PI0\t%.17g\t%.17g\t%.17g\n", c.pi0_, c.left_pi0_, c.right_pi0_);
  dumpVec(f, "CONFIG", c.config_prior_);
  dumpVec(f, "CONFIG_LEFT", c.left_configs_);
  dumpVec(f, "CONFIG_RIGHT", c.right_configs_);
  dumpVec(f, "GRID", c.grid_wts_);
  dumpVec(f, "GRID_LEFT", c.left_grids_);
  dumpVec(f, "GRID_RIGHT", c.right_grids_);
  fprintf(f, "NAMES");
  for (size_t k = 0; k < c.config_names_.size(); ++k) fprintf(f, "\t%s", c.config_names_[k].c_str());
  fprintf(f, "\n");
  if (with_bf) {
    for (size_t g = 0; g < c.genes_.size(); ++g) {
      gene_eQTL &ge = c.genes_[g];
      // (what save_result prints: recomputed with the final parameters, eqtlbma_hm.cpp:1713-1739)
      const double gbf = ge.compute_log10_BF(c.grid_wts_, c.config_prior_, true);
      fprintf(f, "GENE\t%s\t%zu\t%.17g\t%.17g\n", ge.name_.c_str(), ge.snps_.size(), ge.post_prob_gene_, gbf);
      for (size_t p = 0; p < ge.snps_.size(); ++p) {
        fprintf(f, "SNP\t%s\t%.17g\t%.17g", ge.snps_[p].name_.c_str(),
                ge.snps_[p].compute_log10_BF(c.grid_wts_, c.config_prior_, true), ge.snps_[p].post_prob_snp_);
        for (size_t k = 0; k < c.dim_; ++k) fprintf(f, "\t%.17g", ge.snps_[p].compute_log10_config_BF(k, c.grid_wts_));
        fprintf(f, "\n");
      }
      fprintf(f, "POSTCFG");
      for (size_t k = 0; k < ge.post_prob_config_.size(); ++k) fprintf(f, "\t%.17g", ge.post_prob_config_[k].prob);
      fprintf(f, "\n");
    }
  }
  fclose(f);
}

int main(int argc, char **argv)
{
  int verbose = 1, nb_threads = 1;
  string file_pattern, model = "configs", out_file, file_init, file_ci;
  size_t nb_subgroups = string::npos, dim = string::npos, nb_grid_points = string::npos, seed = string::npos,
         max_nb_iters = string::npos;
  double thresh = 0.05, stepmax = 1.0, fixed_pi0 = NaN;
  vector<string> configs_tokeep;
  bool rand_init = false, keep_gen_abfs = false, skip_ci = true, skip_bf = true;
  parseCmdLine(argc, argv, file_pattern, nb_subgroups, model, dim, nb_grid_points, out_file, file_init, rand_init, seed,
               thresh, max_nb_iters, stepmax, nb_threads, file_ci, configs_tokeep, keep_gen_abfs, skip_ci, skip_bf,
               fixed_pi0, verbose);
  if (model != "configs") {
    cerr << "ref_hm_dump: only --model configs is dumped" << endl;
    return EXIT_FAILURE;
  }
  Controller controller(nb_subgroups, model, nb_grid_points, dim, thresh, max_nb_iters, stepmax, fixed_pi0, nb_threads,
                        verbose);
  controller.load_data(file_pattern, configs_tokeep, keep_gen_abfs);
  if (!file_ci.empty()) {
    controller.init_params(file_ci);
    controller.compute_posterior();
    controller.estimate_profile_ci();
    controller.save_result(out_file, true);
    dumpAll(controller, false);
  } else {
    if (!file_init.empty())
      controller.init_params(file_init);
    else
      controller.init_params(seed);
    controller.run_EM();
    controller.compute_log10_ICL();
    if (!skip_bf) controller.compute_posterior();
    if (!skip_ci) controller.estimate_profile_ci();
    controller.save_result(out_file, skip_bf);
    dumpAll(controller, !skip_bf);
  }
  return EXIT_SUCCESS;
}
