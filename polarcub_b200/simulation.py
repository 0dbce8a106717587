"""Monte-Carlo drivers of the reference, batched: encodeDecodeSimulation (BinaryPolarEncoderDecoder.py:328-387),
genieEncodeDecodeSimulation (:390-491), frozenSetFromTVAndPe (:519-548) and the frozen-bits file format (:471-489,
main_deletion.py:149-159).

Same signatures and callbacks as the reference (`make_xVectorDistribution()`, `make_codeword(encodedVector)`,
`simulateChannel(codeword)`, `make_xyVectrorDistribution(receivedWord)`), same random streams (CPython `random`), same
printed summary lines and same frozen-set file -- but the encoder and the decoder run once per BATCH on the GPU: the
callbacks are evaluated trial by trial on the host in the reference's order (they own the channel's RNG state), the
resulting channel outputs are stacked, and one C-ABI call decodes them all.  Uniform a-priori distributions only (the CUDA
path raises otherwise; there is no CPU fallback).
"""
import math
import random
import sys

import numpy as np

from .CollectionOfBinaryTrellises import CollectionOfBinaryTrellises


def eta(p):
    """ScalarDistributions/BinaryMemorylessDistribution.py:451-459"""
    assert 0.0 <= p <= 1.0 + 10 * sys.float_info.epsilon
    p = min(1.0, p)
    return 0.0 if p == 0.0 else -p * math.log2(p)


def frozenSetFromTVAndPe(TVvec, Pevec, errorUpperBoundForFrozenSet, verbose=True):
    """BinaryPolarEncoderDecoder.py:519-548 (stable sort by TV + Pe, greedy error budget)."""
    s = [TVvec[i] + Pevec[i] for i in range(len(TVvec))]
    order = sorted(range(len(s)), key=lambda k: s[k])
    err, idx = 0.0, -1
    while err < errorUpperBoundForFrozenSet and idx + 1 < len(s):
        i = order[idx + 1]
        if s[i] + err <= errorUpperBoundForFrozenSet:
            err += s[i]
            idx += 1
        else:
            break
    frozenSet = set(order[j] for j in range(idx + 1, len(s)))
    if verbose:
        print("frozen set =", frozenSet)
        print("fraction of non-frozen indices =", 1.0 - len(frozenSet) / len(s))
    return frozenSet


def readFrozenSetFromFile(filename):
    """main_deletion.py:149-159"""
    frozenSet = set()
    with open(filename, "r") as f:
        for line in f:
            if line[0] != "*":
                frozenSet.add(int(line))
    return frozenSet


def stack_channel_outputs(length, xys):
    """A list of per-trial xyVectorDistributions -> one batch: trellis-collection descriptors are merged (common sub-word
    width), memoryless distributions (`.probs` [N, 2]) are stacked."""
    if isinstance(xys[0], CollectionOfBinaryTrellises):
        c0 = xys[0]
        maxlen = max(c.sub_bits.shape[2] for c in xys)
        B = sum(c.frames for c in xys)
        bits = np.zeros((B, c0.numberOfTrellises, maxlen), dtype=np.uint8)
        lens = np.zeros((B, c0.numberOfTrellises), dtype=np.int32)
        r = 0
        for c in xys:
            assert (c.n, c.n0, c.ones, c.deletionProb) == (c0.n, c0.n0, c0.ones, c0.deletionProb)
            bits[r:r + c.frames, :, :c.sub_bits.shape[2]] = c.sub_bits
            lens[r:r + c.frames] = c.sub_len
            r += c.frames
        return CollectionOfBinaryTrellises(bits, lens, c0.deletionProb, c0.n, c0.n0, c0.ones)
    return np.stack([np.asarray(getattr(x, "probs", x), dtype=np.float64).reshape(length, 2) for x in xys])


def encodeDecodeSimulation(length, make_xVectorDistribution, make_codeword, simulateChannel, make_xyVectrorDistribution,
                           numberOfTrials, frozenSet, commonRandomnessSeed=1, randomInformationSeed=1, verbosity=0):
    """BinaryPolarEncoderDecoder.py:328-387.  Returns the number of misdecoded words (the reference only prints it)."""
    from .BinaryPolarEncoderDecoder import BinaryPolarEncoderDecoder
    xVectorDistribution = make_xVectorDistribution()
    encDec = BinaryPolarEncoderDecoder(length, frozenSet, commonRandomnessSeed)
    informationRNG = random.Random()
    informationRNG.seed(randomInformationSeed)
    info = np.empty((numberOfTrials, encDec.k), dtype=np.int64)
    for t in range(numberOfTrials):
        for i in range(encDec.k):
            info[t, i] = 0 if informationRNG.random() < 0.5 else 1
    encoded = encDec.encode_batch(info, xVectorDistribution)
    xys = []
    for t in range(numberOfTrials):  # the callbacks own the channel's random state: same order as the reference
        xys.append(make_xyVectrorDistribution(simulateChannel(make_codeword(encoded[t]))))
    batch = stack_channel_outputs(length, xys)
    if isinstance(batch, CollectionOfBinaryTrellises):
        _, decoded = encDec.decode_trellis_batch(batch)
    else:
        _, decoded = encDec.decode_batch(batch, xVectorDistribution)
    bad = (decoded != info).any(axis=1)
    misdecodedWords = int(bad.sum())
    if verbosity > 0:
        for t in np.nonzero(bad)[0]:
            print(str(t) + ") error, transmitted inforamtion:\n" + str(list(info[t])) + "\ndecoded information:\n" + str(decoded[t]))
    print("Error probability = ", misdecodedWords, "/", numberOfTrials, " = ", misdecodedWords / numberOfTrials)
    return misdecodedWords


def genieEncodeDecodeSimulation(length, make_xVectorDistribution, make_codeword, simulateChannel,
                                make_xyVectrorDistribution, numberOfTrials, errorUpperBoundForFrozenSet, genieSeed,
                                trustXYProbs=True, filename=None, return_stats=False):
    """BinaryPolarEncoderDecoder.py:390-491: genie encoder / decoder runs -> TV, Pe, H per index -> frozen set (+ file)."""
    from .BinaryPolarEncoderDecoder import BinaryPolarEncoderDecoder
    xVectorDistribution = make_xVectorDistribution()
    encDec = BinaryPolarEncoderDecoder(length, set(), 0)
    seedRNG = random.Random()
    seedRNG.seed(genieSeed)
    seeds = [seedRNG.randint(1, 1000000) for _ in range(numberOfTrials)]
    encoded, TV, Henc = encDec.genie_encode_batch(xVectorDistribution, seeds)
    xys, codeword = [], None
    for t in range(numberOfTrials):
        codeword = make_codeword(encoded[t])
        xys.append(make_xyVectrorDistribution(simulateChannel(codeword)))
    batch = stack_channel_outputs(length, xys)
    _, Pe, Hdec = encDec.genie_decode_batch(xVectorDistribution, batch, seeds, trustXYProbs)
    # running sums in trial order, exactly as the reference accumulates them (:432-445)
    TVvec, Pevec, HEncvec = TV[0].copy(), Pe[0].copy(), Henc[0].copy()
    HDecvec = Hdec[0].copy() if trustXYProbs else None
    for t in range(1, numberOfTrials):
        TVvec += TV[t]
        Pevec += Pe[t]
        HEncvec += Henc[t]
        if trustXYProbs:
            HDecvec += Hdec[t]
    TVvec, Pevec, HEncvec = list(TVvec / numberOfTrials), list(Pevec / numberOfTrials), list(HEncvec / numberOfTrials)
    HEncsum = 0.0
    for v in HEncvec:
        HEncsum += v
    print("TVVec = ", TVvec)
    print("pevec = ", Pevec)
    print("HEncvec = ", HEncvec)
    if trustXYProbs:
        HDecvec = list(HDecvec / numberOfTrials)
        HDecsum = 0.0
        for v in HDecvec:
            HDecsum += v
        print("HDecvec = ", HDecvec)
    print("Normalized HEncsum = ", HEncsum / len(HEncvec))
    if trustXYProbs:
        print("Normalized HDecsum = ", HDecsum / len(HDecvec))
    frozenSet = frozenSetFromTVAndPe(TVvec, Pevec, errorUpperBoundForFrozenSet)
    print("code rate = ", (len(TVvec) - len(frozenSet)) / len(codeword))
    print("codeword length = ", len(codeword))
    if filename is not None:
        with open(filename, "w") as f:
            f.write("* " + ' '.join(sys.argv[:]) + "\n")
            for i in frozenSet:
                f.write(str(i))
                f.write("\n")
            f.write("** number of trials = " + str(numberOfTrials) + "\n")
            f.write("* (TotalVariation+errorProbability) * (number of trials)" + "\n")
            for i in range(len(TVvec)):
                f.write("*** " + str(i) + " " + str((TVvec[i] + Pevec[i]) * numberOfTrials) + "\n")
    if return_stats:
        return frozenSet, {"TV": TVvec, "Pe": Pevec, "HEnc": HEncvec, "HDec": HDecvec}
    return frozenSet
