"""Genotype decoders of the front-end other than the custom dose matrix (GPU): VCF (GT hard calls, GT not
always the first FORMAT field, phased and unphased separators, one file per subgroup with SNPs missing from
some files) and IMPUTE (probability triplets), both without --scoord (coordinates come from the genotype
files: data_loader.cpp:570-1010, snp.cpp:118-185).  Outputs are compared with the reference's own text
outputs on the same files (tests/golden/{vcf,impute}_input.text.json.gz)."""
import pytest

from scenarios import CLI_SCENARIOS
from test_cli_dropin import DECODERS, _run_and_compare

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", sorted(DECODERS))
def test_cli_reads_vcf_and_impute_like_the_reference(tmp_path, name):
    _run_and_compare(tmp_path, name, threads=1)


@pytest.mark.parametrize("name", sorted(CLI_SCENARIOS))
def test_cli_loader_filters_like_the_reference(tmp_path, name):
    """--sbgrp / --maf / --snp (eqtlbma_bf.cpp:130,193-198): loader-side filters ahead of the hot path."""
    _run_and_compare(tmp_path, name, threads=1)
