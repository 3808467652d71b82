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
public:
    FullProver(std::string zkeyFileNames[], int size);
    ~FullProver();
    void startProve(std::string input, std::string circuit);
    void abort();
    std::string getStatus();   // {"status":"success","proof":"<json text>","pubData":"<json text>"} etc.
    std::string &getErrString() { return errString; }

private:
    // same state machine as the reference (fullprover.hpp:14)
    enum Status { aborted = -2, busy = -1, failed = 0, success = 1, unverified = 2, uninitialized = 3, initializing = 5, ready = 6 };

    struct Circuit {            // one resident zkey: file mapping, parsed header, GPU prover
        std::unique_ptr<BinFileUtils::BinFile> file;
        std::unique_ptr<ZKeyUtils::Header> header;
        std::unique_ptr<Groth16::Prover<AltBn128::Engine>> prover;
    };
    std::map<std::string, Circuit> circuits;

    std::mutex mtx;
    Status status;
    bool canceled;
    std::string pendingInput, pendingCircuit;       // latest request, not started yet
    std::string executingInput, executingCircuit;   // request being proved by the worker thread
    std::string proof, pubData;                     // compact JSON texts of the last result
    std::string errString;

    void checkPending();            // mtx held: start the worker if idle and a request is pending
    void thread_calculateProve();   // worker: witness (popen) -> prove -> calcFinished
    void calcFinished();
    bool isCanceled();
};
#endif
