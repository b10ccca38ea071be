"""Smith-Waterman (SmithWatermanJavaAligner).  CPU part: the oracle against the reference's known
answers (SmithWatermanAlignerAbstractUnitTest.java:143-281).  GPU part: the CUDA aligner against the oracle, bit-exact
(offsets and CIGARs are integers)."""
import numpy as np
import pytest

from oracle import oracle as O

ORIGINAL_DEFAULT = (3, -1, -4, -3)                      # SmithWatermanAlignmentConstants.java:30
STANDARD_NGS = (25, -50, -110, -6)                      # :38
NEW_SW_PARAMETERS = (200, -150, -260, -11)              # :52
ALIGNMENT_TO_BEST_HAPLOTYPE = (10, -15, -30, -5)        # :68

LONG_REF = (b"ATAGAAAATAGTTTTTGGAAATATGGGTGAAGAGACATCTCCTCTTATGGAAAAAGGGATTCTAGAATTTAACAATAAATATTCCCAACTTTCCCCAAGGCTTTAAAATCTACCTTGAAGGAGCAGCTGATG"
            b"TATTTCTAGAACAGACTTAGGTGTCTTGGTGTGGCCTGTAAAGAGATACTGTCTTTCTCTTTTGAGTGTAAGAGAGAAAGGACAGTCTACTCAATAAAGAGTGCTGGGAAAACTGAATATCCACACACAGAATAATAA"
            b"AACTAGATCCTATCTCTCACCATATACAAAGATCAACTCAAAACAAATTAAAGACCTAAATGTAAGACAAGAAATTATAAAACTACTAGAAAAAAACACAAGGGAAATGCTTCAGGACATTGGC")

# (reference, read, params, strategy, expected offset, expected CIGAR): SmithWatermanAlignerAbstractUnitTest.java
KATS = [
    (b"AAAGGACTGACTG", b"ACTGACTGACTG", ORIGINAL_DEFAULT, O.SW_SOFTCLIP, 1, "12M"),                      # :143-151
    (b"AAAGACTACTG", b"AACGGACACTG", (50, -100, -220, -12), O.SW_SOFTCLIP, 1, "2M2I3M1D4M"),               # :153-168
    (b"AAAGACTACTG", b"AACGGACACTG", (200, -50, -300, -22), O.SW_SOFTCLIP, 0, "11M"),
    (b"AAACCCCC", b"CCCCCGGG", ORIGINAL_DEFAULT, O.SW_SOFTCLIP, 3, "5M3S"),                               # :170-178
    (b"TGTGTGTGTGTGTGACAGAGAGAGAGAGAGAGAGAGAGAGAGAGA", b"ACAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGAGA",
     STANDARD_NGS, O.SW_SOFTCLIP, 14, "31M20S"),                                                            # :180-187
    (b"AAACCCCC", b"CCCCC", ORIGINAL_DEFAULT, O.SW_SOFTCLIP, 3, "5M"),                                     # :240-258
    (b"AAACCCCC", b"CCCCC", ORIGINAL_DEFAULT, O.SW_INDEL, 0, "3D5M"),
    (b"AAACCCCC", b"CCCCC", ORIGINAL_DEFAULT, O.SW_LEADING_INDEL, 0, "3D5M"),
    (b"AAACCCCC", b"CCCCC", ORIGINAL_DEFAULT, O.SW_IGNORE, 3, "5M"),
    (LONG_REF, b"AAAAAAA", ORIGINAL_DEFAULT, O.SW_SOFTCLIP, 359, "7M"),                                    # :269-281
    (LONG_REF, b"AAAAAAA", ORIGINAL_DEFAULT, O.SW_INDEL, 0, "1M358D6M29D"),
    (LONG_REF, b"AAAAAAA", ORIGINAL_DEFAULT, O.SW_LEADING_INDEL, 0, "1M1D6M"),
    (LONG_REF, b"AAAAAAA", ORIGINAL_DEFAULT, O.SW_IGNORE, 359, "7M"),
]

PADDED_REF = (b"GCGTCGCAGTCTTAAGGCCCCGCCTTTTCAGACAGCTTCCGCTGGGCCTGGGCCGCTGCGGGGCGGTCACGGCCCCTTTAAGCCTGAGCCCCGCCCCCTGGCTCCCCGCCCCCTCTTCTCCCCTCCCCCAAGCCAGCACCT"
              b"GGTGCCCCGGCGGGTCGTGCGGCGCGGCGCTCCGCGGTGAGCGCCTGACCCCGAGGGGGCCCGGGGCCGCGTCCCTGGGCCCTCCCCACCCTTGCGGTGGCCTCGCGGGTCCCAGGGGCGGGGCTGGAGCGGCAGCAGGG"
              b"CCGGGGAGATGGGCGGTGGGGAGCGCGGGAGGGACCGGGCCGAGCCGGGGGAAGGGCTCCGGTGACT")
PADDED_HAP = (b"GCGTCGCAGTCTTAAGGCCCCGCCTTTTCAGACAGCTTCCGCTGGGCCTGGGCCGCTGCGGGGCGGTCACGGCCCCTTTAAGCCTGAGCCCCGCCCCCTGGCTCCCCGCCCCCTCTTCTCCCCTCCCCCAAGCCAGCACCT"
              b"GGTGCCCCGGCGGGTCGTGCGGCGCGGCGCTCCGCGGTGAGCGCCTGACCCCGA--GGGCC---------------GGGCCCTCCCCACCCTTGCGGTGGCCTCGCGGGTCCCAGGGGCGGGGCTGGAGCGGCAGCAGGG"
              b"CCGGGGAGATGGGCGGTGGGGAGCGCGGGAGGGACCGGGCCGAGCCGGGGGAAGGGCTCCGGTGACT").replace(b"-", b"")
NOT_PADDED_REF = (b"CTTTAAGCCTGAGCCCCGCCCCCTGGCTCCCCGCCCCCTCTTCTCCCCTCCCCCAAGCCAGCACCTGGTGCCCCGGCGGGTCGTGCGGCGCGGCGCTCCGCGGTGAGCGCCTGACCCCGAGGGGGCCCGGGGCCGCGTCC"
                  b"CTGGGCCCTCCCCACCCTTGCGGTGGCCTCGCGGGTCCCAGGGGCGGGGCTGGAGCGGCAGCAGGGCCGGGGAGATGGGCGGTGGGGAGCGCGGGAGGGA")
NOT_PADDED_HAP = (b"CTTTAAGCCTGAGCCCCGCCCCCTGGCTCCCCGCCCCCTCTTCTCCCCTCCCCCAAGCCAGCACCTGGTGCCCCGGCGGGTCGTGCGGCGCGGCGCTCCGCGGTGAGCGCCTGACCCCGA---------GGGCC------"
                  b"--GGGCCCTCCCCACCCTTGCGGTGGCCTCGCGGGTCCCAGGGGCGGGGCTGGAGCGGCAGCAGGGCCGGGGAGATGGGCGGTGGGGAGCGCGGGAGGGA").replace(b"-", b"")


def _non_match_elements(cigar):
    import re
    return [(int(n), op) for n, op in re.findall(r"(\d+)([MIDS])", cigar) if op != "M"]


def test_oracle_known_answers():
    for ref, read, params, strategy, want_off, want_cigar in KATS:
        assert O.sw_align(ref, read, params, strategy) == (want_off, want_cigar), (read, strategy)


def test_oracle_identical_alignments_with_differing_flank_lengths():
    # SmithWatermanAlignerAbstractUnitTest.java:189-238: the indels are placed the same way whatever the flanks
    pad = b"N" * 10
    a = O.sw_align(pad + PADDED_REF + pad, pad + PADDED_HAP + pad, NEW_SW_PARAMETERS, O.SW_SOFTCLIP)
    b = O.sw_align(pad + NOT_PADDED_REF + pad, pad + NOT_PADDED_HAP + pad, NEW_SW_PARAMETERS, O.SW_SOFTCLIP)
    assert _non_match_elements(a[1]) == _non_match_elements(b[1]) and len(_non_match_elements(a[1])) >= 2


def test_oracle_rejects_empty_sequences():
    with pytest.raises(ValueError):
        O.sw_align(b"", b"ACGT", ORIGINAL_DEFAULT, O.SW_SOFTCLIP)    # SmithWatermanJavaAligner.java:64-66


# ---- device ---------------------------------------------------------------------------------------------------------
def _mutated(rng, seq, n_events):
    s = bytearray(seq)
    for _ in range(n_events):
        if len(s) < 4:
            break
        pos = int(rng.integers(0, len(s)))
        kind = int(rng.integers(0, 3))
        if kind == 0:
            s[pos] = b"ACGT"[int(rng.integers(0, 4))]
        elif kind == 1:
            s[pos:pos] = bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 12))).astype(np.uint8))
        else:
            del s[pos:pos + int(rng.integers(1, 12))]
    return bytes(s) if len(s) else b"A"


def _random_pairs(seed, n, ref_len, alt_len):
    rng = np.random.default_rng(seed)
    refs, alts = [], []
    for _ in range(n):
        nr = int(rng.integers(ref_len[0], ref_len[1] + 1))
        ref = bytes(rng.choice(list(b"ACGT"), nr).astype(np.uint8))
        na = int(rng.integers(alt_len[0], alt_len[1] + 1))
        u = rng.random()
        if u < 0.7 and nr >= 2:
            a = int(rng.integers(0, nr - 1))
            alt = _mutated(rng, ref[a:a + na], int(rng.integers(0, 6)))
        elif u < 0.85:
            alt = bytes(rng.choice(list(b"ACGT"), na).astype(np.uint8))
        else:
            alt = _mutated(rng, ref, int(rng.integers(0, 4)))     # haplotype-to-reference style
        refs.append(ref)
        alts.append(alt)
    return refs, alts


@pytest.mark.gpu
def test_sw_known_answers_on_device():
    from gatk_b200.native import GpuPhmm
    with GpuPhmm() as hmm:
        for ref, read, params, strategy, want_off, want_cigar in KATS:
            assert hmm.sw_align([ref], [read], params, strategy) == [(want_off, want_cigar)], (read, strategy)
        pad = b"N" * 10
        got = hmm.sw_align([pad + PADDED_REF + pad, pad + NOT_PADDED_REF + pad], [pad + PADDED_HAP + pad, pad + NOT_PADDED_HAP + pad],
                           NEW_SW_PARAMETERS, O.SW_SOFTCLIP)
        assert _non_match_elements(got[0][1]) == _non_match_elements(got[1][1])


@pytest.mark.gpu
@pytest.mark.parametrize("strategy", [O.SW_SOFTCLIP, O.SW_INDEL, O.SW_LEADING_INDEL, O.SW_IGNORE])
def test_sw_matches_oracle_bit_for_bit(strategy):
    from gatk_b200.native import GpuPhmm
    with GpuPhmm() as hmm:
        for seed, (ref_len, alt_len, n) in enumerate([((1, 40), (1, 40), 300), ((100, 500), (20, 260), 200), ((400, 1100), (300, 1000), 40)]):
            for params in (ORIGINAL_DEFAULT, STANDARD_NGS, NEW_SW_PARAMETERS, ALIGNMENT_TO_BEST_HAPLOTYPE):
                refs, alts = _random_pairs(100 * seed + strategy, n, ref_len, alt_len)
                got = hmm.sw_align(refs, alts, params, strategy, cigar_capacity=2200)
                want = [O.sw_align(r, a, params, strategy) for r, a in zip(refs, alts)]
                assert got == want


@pytest.mark.gpu
def test_sw_errors_and_capacity():
    from gatk_b200 import native
    from gatk_b200.native import GpuPhmm
    with GpuPhmm() as hmm:
        with pytest.raises(native.GpuPhmmError) as e:
            hmm.sw_align([b"ACGT"], [b""], ORIGINAL_DEFAULT, O.SW_SOFTCLIP)       # SmithWatermanJavaAligner.java:64-66
        assert e.value.code == native.ERR_INVALID_ARG
        with pytest.raises(native.GpuPhmmError) as e:
            hmm.sw_align([b"AAAGACTACTG"], [b"AACGGACACTG"], (50, -100, -220, -12), O.SW_SOFTCLIP, cigar_capacity=3)   # 2M2I3M1D4M needs 5
        assert e.value.code == native.ERR_TOO_LARGE
        assert hmm.sw_align([b"AAAGACTACTG"], [b"AACGGACACTG"], (50, -100, -220, -12), O.SW_SOFTCLIP, cigar_capacity=5) == [(1, "2M2I3M1D4M")]
        assert hmm.sw_align([], [], ORIGINAL_DEFAULT, O.SW_SOFTCLIP) == []


@pytest.mark.gpu
def test_sw_plugin_surface_like_the_reference_unit_test():
    # reads like SmithWatermanAlignerAbstractUnitTest.assertAlignmentMatchesExpected (:260-267) with getAligner() -> CUDA
    from gatk_b200 import smithwaterman as sw
    with sw.getAligner(sw.Implementation.CUDA) as aligner:
        alignment = aligner.align(b"AAAGGACTGACTG", b"ACTGACTGACTG", sw.ORIGINAL_DEFAULT, sw.SWOverhangStrategy.SOFTCLIP)
        assert alignment.getAlignmentOffset() == 1 and alignment.getCigar() == "12M"
        # testIndelsAtStartAndEnd (:170-178)
        alignment = aligner.align(b"AAACCCCC", b"CCCCCGGG", sw.ORIGINAL_DEFAULT, sw.SWOverhangStrategy.SOFTCLIP)
        assert (alignment.getAlignmentOffset(), alignment.getCigar()) == (3, "5M3S")
        # getSubstringMatchLong (:269-281), all four strategies in one batch per strategy
        for strategy, want in ((sw.SWOverhangStrategy.SOFTCLIP, (359, "7M")), (sw.SWOverhangStrategy.INDEL, (0, "1M358D6M29D")),
                               (sw.SWOverhangStrategy.LEADING_INDEL, (0, "1M1D6M")), (sw.SWOverhangStrategy.IGNORE, (359, "7M"))):
            got = aligner.alignBatch([LONG_REF] * 3, [b"AAAAAAA"] * 3, sw.ORIGINAL_DEFAULT, strategy)
            assert [(a.getAlignmentOffset(), a.getCigar()) for a in got] == [want] * 3
        with pytest.raises(ValueError):
            aligner.align(b"", b"ACGT", sw.ORIGINAL_DEFAULT, sw.SWOverhangStrategy.SOFTCLIP)
        # a CIGAR with more than 32 elements: the mirror retries with a larger capacity
        rng = np.random.default_rng(3)
        ref = bytes(rng.choice(list(b"ACGT"), 900).astype(np.uint8))
        alt = bytearray(ref)
        for pos in range(880, 20, -20):
            del alt[pos:pos + 3]
        got = aligner.align(ref, bytes(alt), sw.NEW_SW_PARAMETERS, sw.SWOverhangStrategy.SOFTCLIP)
        assert (got.getAlignmentOffset(), got.getCigar()) == O.sw_align(ref, bytes(alt), sw.NEW_SW_PARAMETERS.as_tuple(), O.SW_SOFTCLIP)
        assert got.getCigar().count("D") > 32
