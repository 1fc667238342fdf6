"""Known-answer fixtures transcribed from the reference's own tests (tables and
expected values are data; citations give the reference file:line)."""
import numpy as np

from tskit_b200.tables import Tables

# --- c/tests/testlib.c:31-52
SINGLE_TREE = dict(
    L=1,
    nodes="1 0\n1 0\n1 0\n1 0\n0 1\n0 2\n0 3\n",
    edges="0 1 4 0,1\n0 1 5 2,3\n0 1 6 4,5\n",
    sites="0.125 0\n0.25 0\n0.5 0\n",
    mutations="0 2 1 -1\n1 4 1 -1\n1 0 0 1\n2 0 1 -1\n2 1 1 -1\n2 2 1 -1\n2 3 1 -1\n")
# --- c/tests/testlib.c:70-97
PAPER = dict(
    L=10,
    nodes="1 0\n1 0\n1 0\n1 0\n0 0.071\n0 0.090\n0 0.170\n0 0.202\n0 0.253\n",
    edges="2 10 4 2\n2 10 4 3\n0 10 5 1\n0 2 5 3\n2 10 5 4\n0 7 6 0,5\n7 10 7 0,5\n0 2 8 2,6\n",
    sites="1 0\n4.5 0\n8.5 0\n",
    mutations="0 2 1\n1 0 1\n2 5 1\n")
# --- c/tests/testlib.c:114-150
NONBINARY = dict(
    L=100,
    nodes="1 0\n" * 8 + "0 0.01\n0 0.068\n0 0.130\n0 0.279\n0 0.405\n",
    edges="0 100 8 0,1,2,3\n0 100 9 6,8\n0 100 10 4\n0 17 10 5\n0 100 10 7\n17 100 11 5,9\n"
          "0 17 12 9\n0 100 12 10\n17 100 12 11\n",
    sites="1 0\n18 0\n",
    mutations="0 2 1\n1 11 1\n")
# --- c/tests/testlib.c:157-189 (unary nodes and a non-sample leaf)
UNARY = dict(
    L=100,
    nodes="1 0\n1 0\n1 0\n1 0\n0 0.071\n0 0.090\n0 0.170\n0 0.202\n0 0.253\n0 0\n",
    edges="2 10 4 2,3\n0 10 5 1\n0 2 5 3\n2 10 5 4\n0 10 5 9\n0 7 6 0,5\n7 10 7 0\n0 2 7 2\n"
          "7 10 7 5\n0 7 8 6\n0 2 8 7\n",
    sites="1.0 0\n4.5 0\n8.5 0\n",
    mutations="0 2 1\n1 6 1\n1 9 0\n2 5 1\n")
# --- c/tests/testlib.c:291-320
INTERNAL_SAMPLE = dict(
    L=10,
    nodes="1 0.0\n1 0.1\n1 0.1\n1 0.2\n0 0.4\n1 0.5\n0 0.7\n0 1.0\n0 1.2\n",
    edges="2 8 4 0\n0 10 4 2\n0 2 4 3\n8 10 4 3\n0 10 5 1,4\n8 10 6 0,5\n0 2 7 0,5\n2 8 8 3,5\n",
    sites="1.0 0\n4.5 0\n8.5 0\n",
    mutations="0 2 1\n1 5 1\n2 5 1\n")
# --- c/tests/testlib.c:333-370
MULTIROOT = dict(
    L=10,
    nodes="1 0.0\n" * 6 + "0 0.1\n0 0.2\n0 0.3\n0 0.4\n0 0.8\n0 0.9\n",
    edges="8 10 6 0,3\n0 8 7 3\n0 10 7 4\n8 10 7 6\n4 10 8 1,2\n0 4 9 2\n0 10 9 7\n4 10 9 8\n"
          "0 4 10 0,1\n4 8 11 0,5\n",
    sites="1.0 0\n2.0 0\n3.0 0\n5.0 0\n6.0 0\n8.0 0\n9.0 0\n",
    mutations="0 10 1\n1 9 1\n2 5 1\n3 11 1\n4 9 1\n5 9 1\n6 5 1\n")
# --- c/tests/testlib.c:373-380
EMPTY = dict(L=10, nodes="1 0.0\n" * 6, edges="", sites=None, mutations=None)
# --- c/tests/testlib.c:392-400 (gaps without any tree)
MISSING = dict(
    L=5,
    nodes="1 0.0\n1 0.0\n1 0.0\n0 1.0\n0 2.0\n",
    edges="1.0 2.0 3 0\n1.0 2.0 3 1\n3.0 4.0 3 1\n3.0 4.0 3 2\n3.0 4.0 4 0\n1.0 2.0 4 2\n"
          "1.0 2.0 4 3\n3.0 4.0 4 3\n",
    sites=None, mutations=None)
# --- python/tests/test_tree_stats.py:6230-6276 (SpecificTreesTestCase.test_case_1)
CASE_1 = dict(
    L=1.0,
    nodes="1 0\n1 0\n1 0\n0 0.4\n0 0.5\n0 0.7\n0 1.0\n",
    edges="0.2 0.8 3 0,2\n0.0 0.2 4 1,2\n0.2 0.8 4 1,3\n0.8 1.0 4 1,2\n0.8 1.0 5 0,4\n0.0 0.2 6 0,4\n",
    sites="0.05 0\n0.1 0\n0.11 0\n0.15 0\n0.151 0\n0.3 0\n0.6 0\n0.9 0\n0.95 0\n0.951 0\n",
    mutations="0 4 1\n1 0 1\n2 2 1\n3 0 1\n4 1 1\n5 1 1\n6 2 1\n7 0 1\n8 1 1\n9 2 1\n")
# --- python/tests/test_tree_stats.py:807-859 (four_taxa_test_case)
FOUR_TAXA = dict(
    L=2.5,
    nodes="1 0\n1 0\n1 0\n1 0\n0 0.4\n0 0.5\n0 0.7\n0 1.0\n0 0.4\n",
    edges="0.0 2.5 8 1,3\n0.2 0.8 4 0,2\n0.0 0.2 5 8,2\n0.2 0.8 5 8,4\n0.8 2.5 5 8,2\n"
          "0.8 2.5 6 0,5\n0.0 0.2 7 0,5\n",
    sites=None, mutations=None)

ALL = dict(single_tree=SINGLE_TREE, paper=PAPER, nonbinary=NONBINARY, unary=UNARY,
           internal_sample=INTERNAL_SAMPLE, multiroot=MULTIROOT, empty=EMPTY, missing=MISSING,
           case_1=CASE_1, four_taxa=FOUR_TAXA)


def load(name):
    d = ALL[name]
    return Tables.from_text(d["nodes"], d["edges"], d.get("sites"), d.get("mutations"),
                            sequence_length=d["L"]).ensure_derived()


# ---- expected values -------------------------------------------------------
# c/tests/test_stats.c:1459-1460 (single_tree_ex, all samples as singleton sets)
SINGLE_TREE_D_BRANCH = np.array([0, 2, 6, 6, 2, 0, 6, 6, 6, 6, 0, 4, 6, 6, 4, 0.0]).reshape(4, 4)
SINGLE_TREE_D_SITE = np.array([0, 1, 1, 0, 1, 0, 2, 1, 1, 2, 0, 1, 0, 1, 1, 0.0]).reshape(4, 4)
# c/tests/test_genotypes.c:447-503 (single_tree_ex genotypes per site)
SINGLE_TREE_GENOTYPES = np.array([[0, 0, 1, 0], [0, 1, 0, 0], [1, 1, 1, 1]], dtype=np.int32)
# c/tests/test_stats.c:1718-1743
PAPER_SITE_DIVERSITY = 1.5
# python/tests/test_tree_stats.py:6131-6141 (case 1)
CASE_1_BRANCH_DIVERGENCE = {
    (0, 1): 2 * (1 * (0.2 - 0) + 0.5 * (0.8 - 0.2) + 0.7 * (1.0 - 0.8)),
    (0, 2): 2 * (1 * (0.2 - 0) + 0.4 * (0.8 - 0.2) + 0.7 * (1.0 - 0.8)),
    (1, 2): 2 * (0.5 * (0.2 - 0) + 0.5 * (0.8 - 0.2) + 0.5 * (1.0 - 0.8)),
}
CASE_1_BRANCH_Y3 = 0.2 * (1 + 0.5) + 0.6 * (0.4) + 0.2 * (0.7 + 0.2)
CASE_1_SITE_Y3 = 3 + 0 + 1
# python/tests/test_tree_stats.py:6498-6530 (four taxa)
FOUR_TAXA_F4_0123 = (0.1 * 0.2 + (0.1 + 0.1) * 0.6 + 0.1 * 1.7) / 2.5
FOUR_TAXA_WINDOWS = [0.0, 0.4, 2.5]
FOUR_TAXA_F4_0123_WINDOWED = np.array([(0.1 * 0.2 + (0.1 + 0.1) * 0.2) / 0.4,
                                       ((0.1 + 0.1) * 0.4 + 0.1 * 1.7) / 2.1])
FOUR_TAXA_F2_02_13 = FOUR_TAXA_F4_0123
FOUR_TAXA_DIVERSITY_WINDOWED = (2 / 6) * np.array([
    (0.2 * (1 + 1 + 1 + 0.5 + 0.4 + 0.5) + (0.4 - 0.2) * (0.5 + 0.4 + 0.5 + 0.5 + 0.4 + 0.5)) / 0.4,
    ((0.8 - 0.4) * (0.5 + 0.4 + 0.5 + 0.5 + 0.4 + 0.5)
     + (2.5 - 0.8) * (0.7 + 0.7 + 0.7 + 0.5 + 0.4 + 0.5)) / (2.5 - 0.4)])
