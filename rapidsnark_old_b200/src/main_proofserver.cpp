// proverServer <port> <circuit1.zkey> <circuit2.zkey> ... <circuitN.zkey>
// The reference's proof server (src/main_proofserver.cpp:11-45, src/proverapi.cpp:9-41) over this repository's
// FullProver: the same command line, the same routes and the same status document -
//     GET  /status           -> 200, application/json: FullProver::getStatus()
//     POST /input/:circuit   -> 200: FullProver::startProve(request body, circuit)
//     POST /cancel           -> 200: FullProver::abort()
//     POST /start, /stop     -> 200 (no-ops in the reference too)
// - on plain POSIX sockets instead of pistache (not in this image): HTTP/1.1, one connection at a time like the
// reference's `threads(1)`, bodies up to the reference's maxRequestSize (128 MB), Content-Length framing.  The zkeys
// stay resident on the GPU between requests; proving runs on FullProver's worker thread, as in the reference.
#include <arpa/inet.h>
#include <netinet/in.h>
#include <signal.h>
#include <string.h>
#include <sys/socket.h>
#include <unistd.h>
#include <iostream>
#include <string>
#include <vector>
#include "fullprover.hpp"

static const size_t kMaxRequest = 128000000;

static bool sendAll(int fd, const std::string &s) {
    size_t off = 0;
    while (off < s.size()) {
        ssize_t k = send(fd, s.data() + off, s.size() - off, MSG_NOSIGNAL);
        if (k <= 0) return false;
        off += (size_t)k;
    }
    return true;
}

static void respond(int fd, int code, const char *reason, const std::string &body, const char *type) {
    std::string h = "HTTP/1.1 " + std::to_string(code) + " " + reason + "\r\n";
    if (type) h += std::string("Content-Type: ") + type + "\r\n";
    h += "Content-Length: " + std::to_string(body.size()) + "\r\nConnection: close\r\n\r\n";
    sendAll(fd, h + body);
}

// one request: request line, headers, Content-Length bytes of body
static bool readRequest(int fd, std::string &method, std::string &path, std::string &body) {
    std::string buf;
    char tmp[65536];
    size_t hdrEnd = std::string::npos;
    while (hdrEnd == std::string::npos) {
        ssize_t k = recv(fd, tmp, sizeof tmp, 0);
        if (k <= 0) return false;
        buf.append(tmp, (size_t)k);
        hdrEnd = buf.find("\r\n\r\n");
        if (buf.size() > 65536 && hdrEnd == std::string::npos) return false;
    }
    size_t sp1 = buf.find(' '), sp2 = buf.find(' ', sp1 + 1);
    if (sp1 == std::string::npos || sp2 == std::string::npos || sp2 > hdrEnd) return false;
    method = buf.substr(0, sp1);
    path = buf.substr(sp1 + 1, sp2 - sp1 - 1);
    size_t len = 0;
    {
        std::string head = buf.substr(0, hdrEnd);
        for (char &ch : head) ch = (char)tolower((unsigned char)ch);
        size_t p = head.find("content-length:");
        if (p != std::string::npos) len = (size_t)strtoull(head.c_str() + p + 15, nullptr, 10);
        if (head.find("expect: 100-continue") != std::string::npos) sendAll(fd, "HTTP/1.1 100 Continue\r\n\r\n");
    }
    if (len > kMaxRequest) return false;
    body = buf.substr(hdrEnd + 4);
    while (body.size() < len) {
        ssize_t k = recv(fd, tmp, sizeof tmp, 0);
        if (k <= 0) return false;
        body.append(tmp, (size_t)k);
    }
    body.resize(len);
    return true;
}

int main(int argc, char **argv) {
    if (argc < 3) {
        std::cerr << "Invalid number of parameters:\n";
        std::cerr << "Usage: proverServer <port> <circuit1.zkey> <circuit2.zkey> ... <circuitN.zkey> \n";
        return -1;
    }
    signal(SIGPIPE, SIG_IGN);
    try {
        int port = std::stoi(argv[1]);
        std::vector<std::string> zkeyFileNames(argv + 2, argv + argc);
        std::cerr << "Initializing server...\n";
        FullProver fullProver(zkeyFileNames.data(), (int)zkeyFileNames.size());

        int srv = socket(AF_INET, SOCK_STREAM, 0);
        if (srv < 0) throw std::runtime_error("socket");
        int one = 1;
        setsockopt(srv, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
        sockaddr_in addr;
        memset(&addr, 0, sizeof addr);
        addr.sin_family = AF_INET;
        addr.sin_addr.s_addr = htonl(INADDR_ANY);
        addr.sin_port = htons((uint16_t)port);
        if (bind(srv, (sockaddr *)&addr, sizeof addr) != 0) throw std::runtime_error(std::string("bind: ") + strerror(errno));
        if (listen(srv, 16) != 0) throw std::runtime_error("listen");
        std::cerr << "Server ready on port " << port << "...\n";
        for (;;) {
            int fd = accept(srv, nullptr, nullptr);
            if (fd < 0) continue;
            std::string method, path, body;
            if (!readRequest(fd, method, path, body)) {
                respond(fd, 400, "Bad Request", "", nullptr);
            } else if (method == "GET" && path == "/status") {
                respond(fd, 200, "OK", fullProver.getStatus(), "application/json");
            } else if (method == "POST" && path.compare(0, 7, "/input/") == 0 && path.size() > 7) {
                fullProver.startProve(body, path.substr(7));
                respond(fd, 200, "OK", "", nullptr);
            } else if (method == "POST" && path == "/cancel") {
                fullProver.abort();
                respond(fd, 200, "OK", "", nullptr);
            } else if (method == "POST" && (path == "/start" || path == "/stop")) {
                respond(fd, 200, "OK", "", nullptr);
            } else if (method == "POST" && path == "/quit" && getenv("B200_SERVER_ALLOW_QUIT")) {   // tests only
                respond(fd, 200, "OK", "", nullptr);
                close(fd);
                break;
            } else {
                respond(fd, 404, "Not Found", "", nullptr);
            }
            close(fd);
        }
        close(srv);
    } catch (std::exception &e) {
        std::cerr << e.what() << '\n';
        return -1;
    }
    return 0;
}
