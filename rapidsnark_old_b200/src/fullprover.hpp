// FullProver - the reference's long-lived prover object (src/fullprover.hpp:13-50, fullprover.cpp:21-240):
// loads N zkeys once (here: uploads them to HBM once), runs one proof at a time on a worker thread, witness
// produced by the external circom binary ./build/<circuit>, status machine ready/busy/success/failed/aborted.
// getStatus() returns the same JSON document as the reference, as text.
#ifndef B200_FULLPROVER_HPP
#define B200_FULLPROVER_HPP
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include "binfile_utils.hpp"
#include "groth16.hpp"
#include "zkey_utils.hpp"

class FullProver {
    enum Status { aborted = -2, busy = -1, failed = 0, success = 1, unverified = 2, uninitialized = 3, initializing = 5, ready = 6 };
    Status status;
    std::mutex mtx;

    std::string pendingInput, executingInput, pendingCircuit, executingCircuit;
    std::map<std::string, std::unique_ptr<Groth16::Prover<AltBn128::Engine>>> provers;
    std::map<std::string, std::unique_ptr<ZKeyUtils::Header>> zkHeaders;
    std::map<std::string, std::unique_ptr<BinFileUtils::BinFile>> zKeys;

    std::string proof;    // compact JSON text
    std::string pubData;  // compact JSON text
    std::string errString;
    bool canceled;

    bool isCanceled();
    void calcFinished();
    void thread_calculateProve();
    void checkPending();   // mtx held

public:
    FullProver(std::string zkeyFileNames[], int size);
    ~FullProver();
    void startProve(std::string input, std::string circuit);
    void abort();
    std::string getStatus();   // {"status":"success","proof":"<json text>","pubData":"<json text>"} etc.
    std::string &getErrString() { return errString; }
};
#endif
