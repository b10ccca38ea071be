"""Pure-Python (float64) restatement of LoglessPairHMM for SMALL cases -- an independent cross-check of
oracle/pairhmm_oracle.c.  TEST INFRASTRUCTURE ONLY.

Follows src/main/java/org/broadinstitute/hellbender/utils/pairhmm/LoglessPairHMM.java:20-93 and
PairHMMModel.java:86-117,373-388; the matchToMatch value is taken in closed form
max(0, 1 - eps_i - eps_d), which PairHMMModelUnitTest.java:237-245 pins to 1e-9 of the table.
"""
import math

INITIAL_CONDITION = 2.0 ** 1020
INITIAL_CONDITION_LOG10 = math.log10(INITIAL_CONDITION)


def qual_to_error_prob(q):
    return 10.0 ** (q / -10.0)


def logless(hap, read, base_q, ins_q, del_q, gcp, tristate_off=False):
    H, R = len(hap), len(read)
    Mp = [0.0] * (H + 1)
    Ip = [0.0] * (H + 1)
    Dp = [INITIAL_CONDITION / H] * (H + 1)
    N = ord("N")
    for i in range(1, R + 1):
        ei, ed, eg = (qual_to_error_prob(q[i - 1]) for q in (ins_q, del_q, gcp))
        tMM = max(0.0, 1.0 - ei - ed)
        tIM, tMI, tII, tMD, tDD = 1.0 - eg, ei, eg, ed, eg
        e = qual_to_error_prob(base_q[i - 1])
        p_match, p_mis = 1.0 - e, e / (1.0 if tristate_off else 3.0)
        x = read[i - 1]
        Mc = [0.0] * (H + 1)
        Ic = [0.0] * (H + 1)
        Dc = [0.0] * (H + 1)
        for j in range(1, H + 1):
            y = hap[j - 1]
            prior = p_match if (x == y or x == N or y == N) else p_mis
            Mc[j] = prior * (Mp[j - 1] * tMM + Ip[j - 1] * tIM + Dp[j - 1] * tIM)
            Ic[j] = Mp[j] * tMI + Ip[j] * tII
            Dc[j] = Mc[j - 1] * tMD + Dc[j - 1] * tDD
        Mp, Ip, Dp = Mc, Ic, Dc
    s = 0.0
    for j in range(1, H + 1):
        s += Mp[j] + Ip[j]
    return (math.log10(s) if s > 0.0 else -math.inf) - INITIAL_CONDITION_LOG10


# ---- second, independent restatement of the tandem-repeat scan (strings instead of index arithmetic) -----------------
def find_number_of_repetitions(unit, test, leading):
    """GATKVariantContextUtils.findNumberOfRepetitions(byte[], byte[], boolean), utils/variant/GATKVariantContextUtils.java:949-957"""
    if len(test) == 0:
        return 0
    n = 0
    if leading:
        while test.startswith(unit):
            n += 1
            test = test[len(unit):]
    else:
        while test.endswith(unit):
            n += 1
            test = test[:len(test) - len(unit)]
    return n


def find_tandem_repeat_units(read, offset, max_unit=8, max_repeat=20):
    """ReadLikelihoodCalculationEngine.findTandemRepeatUnits (tools/walkers/haplotypecaller/ReadLikelihoodCalculationEngine.java:193-253);
    returns (unit, repeat length) like the Java Pair"""
    read = bytes(read)
    best_bw, max_bw = read[offset:offset + 1], 0
    for k in range(1, max_unit + 1):
        if offset + 1 - k < 0:
            break
        max_bw = find_number_of_repetitions(read[offset - k + 1:offset + 1], read[:offset + 1], False)
        if max_bw > 1:
            best_bw = read[offset - k + 1:offset + 1]
            break
    best, max_rl = best_bw, max_bw
    if offset < len(read) - 1:
        best_fw, max_fw = read[offset + 1:offset + 2], 0
        for k in range(1, max_unit + 1):
            if offset + k + 1 > len(read):
                break
            max_fw = find_number_of_repetitions(read[offset + 1:offset + k + 1], read[offset + 1:], True)
            if max_fw > 1:
                best_fw = read[offset + 1:offset + k + 1]
                break
        if best_fw == best_bw:
            max_rl = max_bw + max_fw
        else:
            max_rl = max_fw + find_number_of_repetitions(best_fw, read[:offset + 1], False)
        best = best_fw
    return best, min(max_rl, max_repeat)
