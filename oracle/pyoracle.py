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


# ---- LoglessPDPairHMM, full matrices exactly as the Java code keeps them ----------------------------------------------
PD_SNP, PD_DEL_START, PD_DEL_END, PD_A, PD_C, PD_G, PD_T = 1, 2, 4, 8, 16, 32, 64


def pd_logless(hap, pd, read, base_q, ins_q, del_q, gcp, tristate_off=False):
    """utils/pairhmm/LoglessPDPairHMM.java:34-153 with full (R+1) x (H+1) matrices and the state variable declared
    outside the row loop, like the reference"""
    H, R = len(hap), len(read)
    Z = lambda: [[0.0] * (H + 1) for _ in range(R + 1)]
    M, I, D, bM, bI, bD = Z(), Z(), Z(), Z(), Z(), Z()
    for j in range(H + 1):
        D[0][j] = INITIAL_CONDITION / H
    N = ord("N")
    alt_bit = {ord("A"): PD_A, ord("a"): PD_A, ord("C"): PD_C, ord("c"): PD_C, ord("G"): PD_G, ord("g"): PD_G, ord("T"): PD_T, ord("t"): PD_T}
    state = "NORMAL"
    for i in range(1, R + 1):
        ei, ed, eg = (qual_to_error_prob(q[i - 1]) for q in (ins_q, del_q, gcp))
        tMM, tIM, tMI, tII, tMD, tDD = max(0.0, 1.0 - ei - ed), 1.0 - eg, ei, eg, ed, eg
        e = qual_to_error_prob(base_q[i - 1])
        x = read[i - 1]
        for j in range(1, H + 1):
            y, f = hap[j - 1], pd[j - 1]
            match = x == y or x == N or y == N or ((f & PD_SNP) != 0 and (f & alt_bit[x]) != 0)
            prior = (1.0 - e) if match else e / (1.0 if tristate_off else 3.0)
            mx = lambda a, b: a if a > b else b
            if state == "NORMAL":
                bM[i][j], bD[i][j], bI[i][j] = M[i][j - 1], D[i][j - 1], I[i][j - 1]
            elif state == "INSIDE_DEL":
                bM[i][j], bD[i][j], bI[i][j] = bM[i][j - 1], bD[i][j - 1], bI[i][j - 1]
            else:
                bM[i][j], bD[i][j], bI[i][j] = mx(bM[i][j - 1], M[i][j - 1]), mx(bD[i][j - 1], D[i][j - 1]), mx(bI[i][j - 1], I[i][j - 1])
            if state == "AFTER_DEL":
                M[i][j] = prior * (mx(bM[i - 1][j - 1], M[i - 1][j - 1]) * tMM + mx(bI[i - 1][j - 1], I[i - 1][j - 1]) * tIM
                                   + mx(bD[i - 1][j - 1], D[i - 1][j - 1]) * tIM)
                D[i][j] = mx(bM[i][j - 1], M[i][j - 1]) * tMD + mx(bD[i][j - 1], D[i][j - 1]) * tDD
            else:
                M[i][j] = prior * (M[i - 1][j - 1] * tMM + I[i - 1][j - 1] * tIM + D[i - 1][j - 1] * tIM)
                D[i][j] = M[i][j - 1] * tMD + D[i][j - 1] * tDD
            if f & PD_DEL_END:
                I[i][j] = mx(bM[i - 1][j], M[i - 1][j]) * tMI + mx(bI[i - 1][j], I[i - 1][j]) * tII
            else:
                I[i][j] = M[i - 1][j] * tMI + I[i - 1][j] * tII
            if state == "AFTER_DEL":
                state = "NORMAL"
            if f & PD_DEL_START:
                state = "INSIDE_DEL"
            if f & PD_DEL_END:
                state = "AFTER_DEL"
    s = 0.0
    for j in range(1, H + 1):
        s += M[R][j] + I[R][j]
    return (math.log10(s) if s > 0.0 else -math.inf) - INITIAL_CONDITION_LOG10
