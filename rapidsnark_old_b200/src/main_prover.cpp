// prover <circuit.zkey> <witness.wtns> <proof.json> <public.json>
// Same command line, outputs and error behaviour as the reference CLI (src/main_prover.cpp:23-103).
// Extras, all optional and via the environment: B200_GPUS=N (shard the MSMs over N GPUs of this box),
// B200_DEVICE=i (first GPU), B200_TIMING=1 (phase timings to stderr), B200_R / B200_S (hex, test only:
// fixed blinding factors), B200_DUMP_MSMS=file (the five pre-blinding MSM results, 768 bytes).
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include "binfile_utils.hpp"
#include "groth16.hpp"
#include "wtns_utils.hpp"
#include "zkey_utils.hpp"

static bool hexTo32(const char *hex, uint8_t out[32]) {   // big-endian hex string -> little-endian bytes
    size_t n = strlen(hex);
    if (n == 0 || n > 64) return false;
    memset(out, 0, 32);
    for (size_t i = 0; i < n; i++) {
        char ch = hex[n - 1 - i];
        int v = (ch >= '0' && ch <= '9') ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : (ch >= 'A' && ch <= 'F') ? ch - 'A' + 10 : -1;
        if (v < 0) return false;
        out[i / 2] |= (uint8_t)(v << (4 * (i & 1)));
    }
    return true;
}

int main(int argc, char **argv) {
    if (argc != 5) {
        std::cerr << "Invalid number of parameters:\n";
        std::cerr << "Usage: prover <circuit.zkey> <witness.wtns> <proof.json> <public.json>\n";
        return -1;
    }
    typedef std::chrono::steady_clock clk;
    const bool timing = getenv("B200_TIMING") != nullptr;
    try {
        std::string zkeyFilename = argv[1], wtnsFilename = argv[2], proofFilename = argv[3], publicFilename = argv[4];
        auto t0 = clk::now();
        auto zkey = BinFileUtils::openExisting(zkeyFilename, "zkey", 1);
        auto zkeyHeader = ZKeyUtils::loadHeader(zkey.get());
        if (zkeyHeader->n8r != 32 || memcmp(zkeyHeader->rPrime.data(), AltBn128::kFrPrime, 32) != 0)
            throw std::invalid_argument("zkey curve not supported");
        auto wtns = BinFileUtils::openExisting(wtnsFilename, "wtns", 2);
        auto wtnsHeader = WtnsUtils::loadHeader(wtns.get());
        if (wtnsHeader->n8 != 32 || memcmp(wtnsHeader->prime.data(), AltBn128::kFrPrime, 32) != 0)
            throw std::invalid_argument("different wtns curve");
        auto t1 = clk::now();
        setenv("B200_PRECOMP", "0", 0);   // one proof per process: building per-window tables would cost more than it saves
        auto prover = Groth16::makeProver<AltBn128::Engine>(
            zkeyHeader->nVars, zkeyHeader->nPublic, zkeyHeader->domainSize, zkeyHeader->nCoefs, zkeyHeader->vk_alpha1,
            zkeyHeader->vk_beta1, zkeyHeader->vk_beta2, zkeyHeader->vk_delta1, zkeyHeader->vk_delta2,
            zkey->getSectionData(4), zkey->getSectionData(5), zkey->getSectionData(6), zkey->getSectionData(7),
            zkey->getSectionData(8), zkey->getSectionData(9));
        auto t2 = clk::now();
        uint8_t r[32], s[32];
        const char *er = getenv("B200_R"), *es = getenv("B200_S");
        if (er && es) {
            if (!hexTo32(er, r) || !hexTo32(es, s)) throw std::invalid_argument("B200_R / B200_S must be hex");
            prover->setBlinding(r, s);
        }
        AltBn128::FrElement *wtnsData = (AltBn128::FrElement *)wtns->getSectionData(2);
        if (wtns->getSectionSize(2) < (uint64_t)zkeyHeader->nVars * 32) throw std::invalid_argument("witness too short for this zkey");
        auto proof = prover->prove(wtnsData);
        auto t3 = clk::now();

        std::ofstream proofFile;
        proofFile.open(proofFilename);
        proofFile << proof->toJson();
        proofFile.close();

        std::ofstream publicFile;
        publicFile.open(publicFilename);
        // `json jsonPublic;` with nothing pushed streams as `null` (main_prover.cpp:85-93 with nPublic == 0)
        publicFile << (zkeyHeader->nPublic ? "[" : "null");
        for (uint32_t i = 1; i <= zkeyHeader->nPublic; i++) {
            if (i > 1) publicFile << ",";
            publicFile << "\"" << AltBn128::le32ToString(&wtnsData[i]) << "\"";
        }
        if (zkeyHeader->nPublic) publicFile << "]";
        publicFile.close();

        if (const char *dump = getenv("B200_DUMP_MSMS")) {
            std::ofstream f(dump, std::ios::binary);
            f.write((const char *)prover->lastMsms, 768);
        }
        if (timing) {
            auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
            std::cerr << "open+headers " << ms(t0, t1) << " ms, makeProver(upload) " << ms(t1, t2) << " ms, prove "
                      << ms(t2, t3) << " ms on " << prover->gpuCount() << " GPU(s)\n";
            for (auto &p : prover->lastPhases) std::cerr << "  gpu0 " << p.first << " " << p.second << " ms\n";
        }
    } catch (std::exception &e) {
        std::cerr << e.what() << '\n';
        return -1;
    }
    return 0;
}
