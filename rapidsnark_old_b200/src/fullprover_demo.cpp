// fullprover_demo <circuit.zkey> [<circuit2.zkey> ...] -- <circuit-name> <input.json>
// Drives FullProver the way the reference's REST handlers do (src/proverapi.cpp:9-41): startProve, poll getStatus
// until it leaves "busy", print the final status document.  The witness comes from ./build/<circuit-name>
// (process boundary, src/fullprover.cpp:117-132), exactly like the reference.
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>
#include <vector>
#include "fullprover.hpp"

int main(int argc, char **argv) {
    std::vector<std::string> zkeys;
    int i = 1;
    for (; i < argc && std::string(argv[i]) != "--"; i++) zkeys.push_back(argv[i]);
    if (zkeys.empty() || argc - i != 3) {
        std::cerr << "usage: fullprover_demo <circuit.zkey>... -- <circuit-name> <input.json>\n";
        return 2;
    }
    try {
        FullProver fp(zkeys.data(), (int)zkeys.size());
        std::cout << fp.getStatus() << std::endl;                    // {"status":"ready"}
        std::ifstream in(argv[i + 2]);
        std::stringstream ss;
        ss << in.rdbuf();
        fp.startProve(ss.str(), argv[i + 1]);
        std::string st;
        for (int k = 0; k < 6000; k++) {
            st = fp.getStatus();
            if (st.find("\"busy\"") == std::string::npos) break;
            std::this_thread::sleep_for(std::chrono::milliseconds(5));
        }
        std::cout << st << std::endl;
        return st.find("\"success\"") != std::string::npos ? 0 : 1;
    } catch (std::exception &e) {
        std::cerr << e.what() << '\n';
        return 3;
    }
}
