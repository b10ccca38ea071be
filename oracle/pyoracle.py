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
