#include "fullprover.hpp"
#include <string.h>
#include <array>
#include <fstream>
#include <iostream>
#include <stdexcept>
#include <thread>
#include "minijson.hpp"
#include "wtns_utils.hpp"

static std::string getfilename(std::string path) {
    path = path.substr(path.find_last_of("/\\") + 1);
    size_t dot_i = path.find_last_of('.');
    return path.substr(0, dot_i);
}

static std::string jsonEscape(const std::string &s) {
    std::string o;
    minijson::escape(o, s);
    return o;
}

FullProver::FullProver(std::string zkeyFileNames[], int size) : status(uninitialized), canceled(false) {
    for (int i = 0; i < size; i++) {
        std::string circuit = getfilename(zkeyFileNames[i]);
        Circuit &ck = circuits[circuit];
        ck.file = BinFileUtils::openExisting(zkeyFileNames[i], "zkey", 1);
        ck.header = ZKeyUtils::loadHeader(ck.file.get());
        auto &h = ck.header;
        if (h->n8r != 32 || memcmp(h->rPrime.data(), AltBn128::kFrPrime, 32) != 0)
            throw std::invalid_argument("zkey curve not supported");
        auto &z = ck.file;
        ck.prover = Groth16::makeProver<AltBn128::Engine>(
            h->nVars, h->nPublic, h->domainSize, h->nCoefs, h->vk_alpha1, h->vk_beta1, h->vk_beta2, h->vk_delta1,
            h->vk_delta2, z->getSectionData(4), z->getSectionData(5), z->getSectionData(6), z->getSectionData(7),
            z->getSectionData(8), z->getSectionData(9));
    }
    status = ready;
}

FullProver::~FullProver() {
    // a detached worker may still be running: wait for it (the reference simply races here)
    for (;;) {
        { std::lock_guard<std::mutex> g(mtx); if (status != busy) break; }
        std::this_thread::sleep_for(std::chrono::milliseconds(5));
    }
}

void FullProver::startProve(std::string input, std::string circuit) {
    std::lock_guard<std::mutex> guard(mtx);
    pendingInput = input;
    pendingCircuit = circuit;
    if (status == busy) canceled = true;   // fullprover.cpp:75-77 (without re-locking the mutex)
    checkPending();
}

void FullProver::checkPending() {
    if (status != busy) {
        if (pendingInput != "" && pendingCircuit != "") {
            status = busy;
            executingInput = pendingInput;
            executingCircuit = pendingCircuit;
            pendingInput = "";
            pendingCircuit = "";
            errString = "";
            canceled = false;
            proof = "null";
            std::thread th(&FullProver::thread_calculateProve, this);
            th.detach();
        }
    }
}

void FullProver::thread_calculateProve() {
    try {
        std::string circuit = executingCircuit;
        auto it = circuits.find(circuit);
        if (it == circuits.end()) throw std::runtime_error("unknown circuit: " + circuit);
        Circuit &ck = it->second;
        {
            // fullprover.cpp:108-113: the request body is parsed first (malformed input fails the request here, before
            // the witness generator is spawned) and written back as nlohmann's compact dump
            minijson::Value j = minijson::parse(executingInput);
            std::ofstream file("./build/input_" + circuit + ".json");
            file << minijson::dump(j);
        }
        // witness generation by the circom-generated binary (process boundary, fullprover.cpp:117-132)
        std::string witnessFile("./build/" + circuit + ".wtns");
        std::string command("./build/" + circuit + " ./build/input_" + circuit + ".json " + witnessFile);
        std::array<char, 128> buffer;
        std::string result;
        FILE *pipe = popen(command.c_str(), "r");
        if (!pipe) throw std::runtime_error("Couldn't start command.");
        while (fgets(buffer.data(), 128, pipe) != NULL) result += buffer.data();
        int returnCode = pclose(pipe);
        std::cout << result << std::endl;
        std::cout << returnCode << std::endl;

        std::unique_ptr<BinFileUtils::BinFile> wtns;
        try {
            wtns = BinFileUtils::openExisting(witnessFile, "wtns", 2);
        } catch (std::exception &e) {
            throw std::runtime_error(std::string("witness file: ") + e.what());
        }
        auto wtnsHeader = WtnsUtils::loadHeader(wtns.get());
        if (wtnsHeader->n8 != 32 || memcmp(wtnsHeader->prime.data(), AltBn128::kFrPrime, 32) != 0)
            throw std::runtime_error("different wtns curve");
        AltBn128::FrElement *wtnsData = (AltBn128::FrElement *)wtns->getSectionData(2);
        if (wtns->getSectionSize(2) < (uint64_t)ck.header->nVars * 32) throw std::runtime_error("witness too short for this zkey");

        // a json that nothing was pushed to dumps as `null`, not `[]` (fullprover.cpp:145-150 with nPublic == 0)
        std::string pd = ck.header->nPublic ? "[" : "null";
        for (uint32_t i = 1; i <= ck.header->nPublic; i++) {
            if (i > 1) pd += ",";
            pd += "\"" + AltBn128::le32ToString(&wtnsData[i]) + "\"";
        }
        if (ck.header->nPublic) pd += "]";
        pubData = pd;

        if (!isCanceled()) proof = ck.prover->prove(wtnsData)->toJson();
        else proof = "null";
        calcFinished();
    } catch (std::runtime_error &e) {
        if (!isCanceled()) errString = e.what();
        calcFinished();
    } catch (std::exception &e) {   // the reference would terminate here; report instead
        if (!isCanceled()) errString = e.what();
        calcFinished();
    }
}

void FullProver::calcFinished() {
    std::lock_guard<std::mutex> guard(mtx);
    if (canceled) status = aborted;
    else if (errString != "") status = failed;
    else status = success;
    canceled = false;
    executingInput = "";
    checkPending();
}

bool FullProver::isCanceled() {
    std::lock_guard<std::mutex> guard(mtx);
    return canceled;
}

void FullProver::abort() {
    std::lock_guard<std::mutex> guard(mtx);
    if (status != busy) return;
    canceled = true;
}

std::string FullProver::getStatus() {
    std::lock_guard<std::mutex> guard(mtx);
    if (status == ready) return "{\"status\":\"ready\"}";
    if (status == aborted) return "{\"status\":\"aborted\"}";
    if (status == failed) return "{\"error\":\"" + jsonEscape(errString) + "\",\"status\":\"failed\"}";
    if (status == success)
        return "{\"proof\":\"" + jsonEscape(proof) + "\",\"pubData\":\"" + jsonEscape(pubData) + "\",\"status\":\"success\"}";
    if (status == busy) return "{\"status\":\"busy\"}";
    return "{}";
}
