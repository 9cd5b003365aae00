/*
 * ref_dump_main.cpp -- TEST INFRASTRUCTURE (oracle side), never part of the product path.
 *
 * Drives the UNMODIFIED reference (it textually includes the reference's own
 * src/eqtlbma_bf.cpp from where it lies under /root/reference, with `main` renamed) through
 * the same sequence as the reference's run() (eqtlbma_bf.cpp:1449-1582): load, then per
 * write-group testForAssociations -> makePermutations -> writeRes, and in addition dumps every
 * result the writers would print at FULL precision (%.17g) into the file named by
 * $EQTLBMA_DUMP.  The reference's text outputs only carry 7 significant digits; the dump is
 * what the 1e-9 / 1e-8 parity checks of the oracle restatement are pinned against.
 */
#define main eqtlbma_bf_reference_main
#include REF_BF_CPP
#undef main

static void dumpGroup(FILE *f, const vector<string> &subgroups,
                      map<string, Gene>::iterator itG_begin, map<string, Gene>::iterator itG_end,
                      const string &analysis, const string &bfs, const string &error_model,
                      const Grid &iGridL, const Grid &iGridS, bool is_perm, int perm_sep,
                      const string &permbf, bool use_max_bf, size_t nb_permutations)
{
  for (map<string, Gene>::iterator itG = itG_begin; itG != itG_end; ++itG) {
    Gene &gene = itG->second;
    fprintf(f, "GENE\t%s\t%zu\n", itG->first.c_str(), gene.GetNbGeneSnpPairs());
    for (vector<GeneSnpPair>::const_iterator p = gene.BeginPair(); p != gene.EndPair(); ++p) {
      fprintf(f, "PAIR\t%s\t%s\t%zu\n", itG->first.c_str(), p->GetSnpName().c_str(),
              p->GetNbSubgroups());
      if (error_model != "mvlr") {
        for (size_t s = 0; s < subgroups.size(); ++s) {
          if (!p->HasResults(subgroups[s])) continue;
          fprintf(f, "SS\t%zu\t%zu\t%.17g\t%.17g\t%.17g\t%.17g\t%.17g\n", s,
                  p->GetSampleSize(subgroups[s]), p->GetPve(subgroups[s]),
                  p->GetSigmahat(subgroups[s]), p->GetBetahatGeno(subgroups[s]),
                  p->GetSebetahatGeno(subgroups[s]), p->GetBetapvalGeno(subgroups[s]));
        }
      }
      if (analysis == "join") {
        vector<string> names;
        names.push_back("gen");
        names.push_back("gen-fix");
        names.push_back("gen-maxh");
        if (bfs != "gen") {
          for (size_t k = 1; k <= subgroups.size(); ++k) {
            gsl_combination *comb = gsl_combination_calloc(subgroups.size(), k);
            while (true) {
              stringstream ss;
              ss << gsl_combination_get(comb, 0) + 1;
              for (size_t i = 1; i < k; ++i) ss << "-" << gsl_combination_get(comb, i) + 1;
              names.push_back(ss.str());
              if (gsl_combination_next(comb) != GSL_SUCCESS) break;
            }
            gsl_combination_free(comb);
            if (bfs == "sin") break;
          }
        }
        for (size_t i = 0; i < names.size(); ++i) {
          fprintf(f, "RAW\t%s", names[i].c_str());
          for (vector<double>::const_iterator it = p->BeginUnweightedAbf(names[i]);
               it != p->EndUnweightedAbf(names[i]); ++it)
            fprintf(f, "\t%.17g", *it);
          fprintf(f, "\n");
          fprintf(f, "W\t%s\t%.17g\n", names[i].c_str(), p->GetWeightedAbf(names[i]));
        }
        if (bfs == "sin" || bfs == "all")
          fprintf(f, "W\tgen-sin\t%.17g\n", p->GetWeightedAbf("gen-sin"));
        if (bfs == "all") fprintf(f, "W\tall\t%.17g\n", p->GetWeightedAbf("all"));
      }
    }
    if (is_perm && gene.GetNbGeneSnpPairs() > 0) {
      if (analysis == "join")
        fprintf(f, "PERMJOIN\t%s\t%zu\t%.17g\t%zu\t%.17g\t%.17g\t%zu\n", itG->first.c_str(),
                gene.GetNbGeneSnpPairs(), gene.GetPermutationPvalueJoin(),
                gene.GetNbPermutationsJoin(), gene.GetTrueL10Abf(use_max_bf),
                gene.GetMedianPermL10Abf(), nb_permutations);
      else if (perm_sep == 1)
        fprintf(f, "PERMSEP1\t%s\t%zu\t%.17g\t%zu\t%.17g\t%zu\n", itG->first.c_str(),
                gene.GetNbGeneSnpPairs(), gene.GetPermutationPvalueSep(),
                gene.GetNbPermutationsSep(), gene.GetTrueMinPval(), nb_permutations);
      else if (perm_sep == 2)
        for (size_t s = 0; s < subgroups.size(); ++s)
          fprintf(f, "PERMSEP2\t%s\t%zu\t%zu\t%.17g\t%zu\t%.17g\t%zu\n", itG->first.c_str(), s,
                  gene.GetNbGeneSnpPairs(subgroups[s]), gene.GetPermutationPvalueSep(subgroups[s]),
                  gene.GetNbPermutationsSep(subgroups[s]), gene.GetTrueMinPval(subgroups[s]),
                  nb_permutations);
    }
  }
}

int main(int argc, char **argv)
{
  int verbose = 1, trick = 0, perm_sep = 0, nb_threads = 1, write_group_size = 10;
  size_t radius = 100000, nb_types = string::npos, nb_permutations = 0, seed = string::npos,
         trick_cutoff = 10;
  float min_maf = 0.0, prop_cov_errors = 0.5;
  bool save_sstats = false, save_weighted_abfs = false, need_qnorm = false, use_max_bf = false;
  string file_genopaths, file_snpcoords, file_exppaths, file_genecoords, anchor = "TSS",
         file_sstats, out_prefix, likelihood = "normal", analysis, file_covarpaths,
         file_largegrid, file_smallgrid, bfs = "gen", error_model = "uvlr", permbf = "none",
         file_snpstokeep;
  vector<string> subgroups_tokeep;

  parseCmdLine(argc, argv, file_genopaths, file_snpcoords, file_exppaths, file_genecoords, anchor,
               radius, file_sstats, out_prefix, save_sstats, save_weighted_abfs, likelihood,
               analysis, need_qnorm, min_maf, file_covarpaths, file_largegrid, file_smallgrid, bfs,
               error_model, prop_cov_errors, nb_types, nb_permutations, seed, trick, trick_cutoff,
               perm_sep, permbf, use_max_bf, nb_threads, file_snpstokeep, subgroups_tokeep,
               write_group_size, verbose);

  const char *dump_path = getenv("EQTLBMA_DUMP");
  if (dump_path == NULL) {
    fprintf(stderr, "ERROR: set EQTLBMA_DUMP to the path of the full-precision dump\n");
    return EXIT_FAILURE;
  }
  FILE *f = fopen(dump_path, "w");
  if (f == NULL) {
    perror("EQTLBMA_DUMP");
    return EXIT_FAILURE;
  }

  set<string> sSnpsToKeep;
  if (!file_snpstokeep.empty()) loadSnpsToKeep(file_snpstokeep, verbose, sSnpsToKeep);

  vector<string> subgroups;
  Samples samples;
  map<string, Snp> snp2object;
  map<string, vector<Snp *> > mChr2VecPtSnps;
  Covariates covariates;
  map<string, Gene> gene2object;
  loadRawInputData(file_genopaths, file_snpcoords, file_exppaths, file_genecoords, anchor, radius,
                   min_maf, file_covarpaths, error_model, subgroups_tokeep, sSnpsToKeep, verbose,
                   subgroups, samples, snp2object, mChr2VecPtSnps, covariates, gene2object);

  Grid iGridL(file_largegrid, true, verbose);
  Grid iGridS(file_smallgrid, false, verbose);
  writeRes(out_prefix, save_sstats, save_weighted_abfs, subgroups, gene2object.begin(),
           gene2object.end(), snp2object, analysis, iGridL, iGridS, bfs, nb_permutations, perm_sep,
           error_model, seed, permbf, use_max_bf, "only");

  fprintf(f, "SUBGROUPS");
  for (size_t s = 0; s < subgroups.size(); ++s) fprintf(f, "\t%s", subgroups[s].c_str());
  fprintf(f, "\n");

  bool is_perm = nb_permutations > 0 && (perm_sep != 0 || permbf != "none");
  size_t nbAnalyzedGenes = 0, nbAnalyzedPairs = 0;
  for (map<string, Gene>::iterator itG = gene2object.begin(); itG != gene2object.end();) {
    map<string, Gene>::iterator itG_begin = itG;
    size_t step_size = min((int)distance(itG, gene2object.end()), write_group_size);
    advance(itG, step_size);
    testForAssociations(true, mChr2VecPtSnps, anchor, radius, subgroups, samples, likelihood,
                        analysis, need_qnorm, covariates, iGridL, iGridS, bfs, error_model,
                        prop_cov_errors, verbose, itG_begin, itG, nbAnalyzedGenes, nbAnalyzedPairs);
    if (is_perm) {
      omp_set_num_threads(nb_threads);
      makePermutations(subgroups, samples, likelihood, analysis, need_qnorm, covariates, iGridL,
                       iGridS, error_model, prop_cov_errors, nb_permutations, seed, trick,
                       trick_cutoff, perm_sep, permbf, use_max_bf, itG_begin, itG);
    }
    writeRes(out_prefix, save_sstats, save_weighted_abfs, subgroups, itG_begin, itG, snp2object,
             analysis, iGridL, iGridS, bfs, nb_permutations, perm_sep, error_model, seed, permbf,
             use_max_bf, "none");
    dumpGroup(f, subgroups, itG_begin, itG, analysis, bfs, error_model, iGridL, iGridS, is_perm,
              perm_sep, permbf, use_max_bf, nb_permutations);
    gene2object.erase(itG_begin, itG);
  }
  fprintf(f, "END\t%zu\t%zu\n", nbAnalyzedPairs, nbAnalyzedGenes);
  fclose(f);
  return EXIT_SUCCESS;
}
