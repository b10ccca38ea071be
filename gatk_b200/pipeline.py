"""Caller-side half of the cross-region queue, as a host-side mirror.

The reference calls one region at a time (HaplotypeCallerEngine.callRegion, HC/HaplotypeCallerEngine.java:913-994) and
RampedHaplotypeCallerEngine already splits that call into phases (HC/RampedHaplotypeCallerEngine.java:54-62:
prepare, assemble, computeReadLikelihoods, uncollapse, filter, genotype).  A pipelined caller runs the phases before the
likelihoods for region n+1 .. n+K while the GPU works on region n, and finishes regions strictly in order so that the
VCF writer sees the same sequence (HC/HaplotypeCaller.java:283-285).  RegionPipeline is that schedule over
gphmm_submit_regions / gphmm_wait: `feed` plays prepare+assemble, the ticket is the in-flight computeReadLikelihoods,
`drain` plays uncollapse+filter+genotype.  It holds no arithmetic.
"""
from collections import deque
from typing import Callable, Deque, Iterable, Iterator, Tuple


class RegionPipeline:
    def __init__(self, hmm, lookahead: int = 16, **region_step_params):
        """hmm: gatk_b200.native.GpuPhmm; lookahead: regions in flight (K); region_step_params: see GpuPhmm._region_steps"""
        if lookahead < 1:
            raise ValueError("lookahead must be at least 1")
        self._hmm = hmm
        self._k = lookahead
        self._params = region_step_params
        self._inflight: Deque[Tuple[object, int]] = deque()

    def run(self, regions: Iterable, feed: Callable, drain: Callable) -> Iterator:
        """regions: any iterable; feed(region) -> (batch, mapq, ref_hap) (the phases before the likelihoods);
        drain(region, result_dict) -> value (the phases after).  Yields drain's values in region order."""
        for region in regions:
            batch, mapq, ref_hap = feed(region)
            self._inflight.append((region, self._hmm.submit_regions(batch, mapq, ref_hap, **self._params)))
            if len(self._inflight) >= self._k:
                yield self._finish_oldest(drain)
        while self._inflight:
            yield self._finish_oldest(drain)

    def _finish_oldest(self, drain):
        region, ticket = self._inflight.popleft()
        return drain(region, self._hmm.wait(ticket))
