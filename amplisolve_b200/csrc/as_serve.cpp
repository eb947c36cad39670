// Resident service for the drop-in programs.
//
// A run of AmpliSolveErrorEstimation / AmpliSolveVariantCalling / computeCounts on a gene-panel-sized job does a few tenths
// of a second of work and then waits for the CUDA driver: an EMPTY CUDA process takes 0.25 - 5 s to get its context on the
// B200 boxes (profiles/r02_cuda_startup.txt), once per process, and a job is two or three processes.  The reference has no
// counterpart (it has no GPU); this is the deployment answer to that cost and changes nothing in the programs' interface:
//
//     amplisolve_b200/bin/amplisolve_b200_serve socket=/tmp/as.sock [devices=0,1] &      # holds the CUDA context
//     AS_SERVER=/tmp/as.sock amplisolve_b200/bin/AmpliSolveErrorEstimation panel_design=... (the usual arguments)
//
// With AS_SERVER set, the thin main hands its argv, working directory, AS_* environment and its own stdout / stderr
// descriptors (SCM_RIGHTS) to the service, which runs the very same as_*_main in its process -- same code, same
// outputs, same exit status -- and answers with the status.  No service listening (or any failure before the program
// starts there): the program runs in its own process as always.  One request at a time, in arrival order.
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/un.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "amplisolve_b200.h"

extern char** environ;

namespace {

const uint32_t kMagic = 0x41533142u;  // "AS1B"
bool g_resident = false;

bool write_all(int fd, const void* p, size_t n) {
    const char* c = (const char*)p;
    while (n > 0) {
        const ssize_t w = ::send(fd, c, n, MSG_NOSIGNAL);
        if (w < 0) { if (errno == EINTR) continue; return false; }
        c += w; n -= (size_t)w;
    }
    return true;
}
bool read_all(int fd, void* p, size_t n) {
    char* c = (char*)p;
    while (n > 0) {
        const ssize_t r = ::recv(fd, c, n, 0);
        if (r < 0) { if (errno == EINTR) continue; return false; }
        if (r == 0) return false;
        c += r; n -= (size_t)r;
    }
    return true;
}

int connect_to(const char* path) {
    sockaddr_un a;
    memset(&a, 0, sizeof a);
    a.sun_family = AF_UNIX;
    if (strlen(path) >= sizeof a.sun_path) return -1;
    strcpy(a.sun_path, path);
    const int fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd < 0) return -1;
    if (connect(fd, (sockaddr*)&a, sizeof a) != 0) { close(fd); return -1; }
    return fd;
}

int run_program(uint32_t prog, int argc, char** argv) {
    switch (prog) {
        case 0: return as_error_estimation_main(argc, argv);
        case 1: return as_variant_calling_main(argc, argv);
        case 2: return as_compute_counts_main(argc, argv);
        default: return 2;
    }
}

}  // namespace

// as_host.cpp / as_bam.cpp: in a resident process a program's context is destroyed when the program returns (in its own
// process it is left to _exit)
extern "C" int as_process_is_resident(void) { return g_resident ? 1 : 0; }

// Client side.  Returns 1 when the service ran the program (*rc_out = its exit status), 0 when the caller must run it itself.
extern "C" int as_client_run(int prog, int argc, char** argv, int* rc_out) {
    const char* path = getenv("AS_SERVER");
    if (!path || !*path || g_resident) return 0;
    const int fd = connect_to(path);
    if (fd < 0) return 0;
    std::string body;
    auto put32 = [&](uint32_t v) { body.append((const char*)&v, 4); };
    auto puts_ = [&](const char* s) { body.append(s, strlen(s) + 1); };
    char cwd[4096];
    if (!getcwd(cwd, sizeof cwd)) { close(fd); return 0; }
    std::vector<const char*> env;
    for (char** e = environ; e && *e; ++e)
        if (strncmp(*e, "AS_", 3) == 0 && strncmp(*e, "AS_SERVER=", 10) != 0) env.push_back(*e);
    put32(kMagic); put32((uint32_t)prog); put32((uint32_t)argc); put32((uint32_t)env.size());
    puts_(cwd);
    for (int i = 0; i < argc; ++i) puts_(argv[i]);
    for (const char* e : env) puts_(e);
    std::cout.flush();
    fflush(nullptr);
    // first message: the length, with this process's stdout and stderr as ancillary data
    const uint32_t len = (uint32_t)body.size();
    msghdr m;
    memset(&m, 0, sizeof m);
    iovec io = {(void*)&len, 4};
    m.msg_iov = &io;
    m.msg_iovlen = 1;
    char ctl[CMSG_SPACE(2 * sizeof(int))];
    memset(ctl, 0, sizeof ctl);
    m.msg_control = ctl;
    m.msg_controllen = sizeof ctl;
    cmsghdr* c = CMSG_FIRSTHDR(&m);
    c->cmsg_level = SOL_SOCKET;
    c->cmsg_type = SCM_RIGHTS;
    c->cmsg_len = CMSG_LEN(2 * sizeof(int));
    const int fds[2] = {1, 2};
    memcpy(CMSG_DATA(c), fds, sizeof fds);
    if (sendmsg(fd, &m, MSG_NOSIGNAL) != 4 || !write_all(fd, body.data(), body.size())) { close(fd); return 0; }
    // the service answers "started" before it runs the program: from here on the program is NOT run a second time.  A
    // service that is busy with other clients (or hung) for longer than AS_SERVER_WAIT seconds (default 10) is not waited
    // for: the program runs in its own process instead.
    uint32_t started = 0;
    {
        const char* w = getenv("AS_SERVER_WAIT");
        timeval tv = {w ? atol(w) : 10, 0};
        setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
        if (!read_all(fd, &started, 4) || started != kMagic) { close(fd); return 0; }
        tv.tv_sec = 0;
        setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);  // the program itself may take as long as it takes
    }
    int32_t rc = 1;
    if (!read_all(fd, &rc, 4)) {
        fprintf(stderr, "amplisolve_b200: the service at %s went away while running the program\n", path);
        rc = 1;
    }
    close(fd);
    *rc_out = rc;
    return 1;
}

// amplisolve_b200_serve socket=<path> [devices=0,1,...]
extern "C" int as_serve_main(int argc, char** argv) {
    std::string path, devices;
    for (int i = 1; i < argc; ++i) {
        if (strncmp(argv[i], "socket=", 7) == 0) path = argv[i] + 7;
        else if (strncmp(argv[i], "devices=", 8) == 0) devices = argv[i] + 8;
    }
    if (path.empty()) {
        printf("\nUsage: amplisolve_b200_serve socket=<path of a UNIX socket> [devices=0,1,...]\n"
               "  keeps a CUDA context open and runs the drop-in programs of clients that set AS_SERVER=<path>\n\n");
        return 0;
    }
    // before the first CUDA call; ordinals become 0..n-1.  Without it every GPU of the box stays visible and the programs
    // choose among them per job (the client's AS_DEVICES applies)
    if (!devices.empty()) setenv("CUDA_VISIBLE_DEVICES", devices.c_str(), 1);
    signal(SIGPIPE, SIG_IGN);
    g_resident = true;
    if (!getenv("AS_SERVE_NO_WARMUP")) {  // the context: created once, kept by the runtime for the life of the process
        as_ctx* warm = nullptr;  // (tests of the plumbing on a machine without a GPU skip this)
        if (as_create(0, &warm) != AS_OK) {
            fprintf(stderr, "amplisolve_b200_serve: %s\n", as_last_error());
            return 1;
        }
        as_destroy(warm);
    }
    sockaddr_un a;
    memset(&a, 0, sizeof a);
    a.sun_family = AF_UNIX;
    if (path.size() >= sizeof a.sun_path) { fprintf(stderr, "amplisolve_b200_serve: socket path too long\n"); return 1; }
    strcpy(a.sun_path, path.c_str());
    unlink(path.c_str());
    const int ls = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    const mode_t old = umask(0077);  // the socket is the owner's
    const bool bound = ls >= 0 && bind(ls, (sockaddr*)&a, sizeof a) == 0 && listen(ls, 64) == 0;
    umask(old);
    if (!bound) { fprintf(stderr, "amplisolve_b200_serve: cannot listen on %s: %s\n", path.c_str(), strerror(errno)); return 1; }
    fprintf(stderr, "amplisolve_b200_serve: ready on %s\n", path.c_str());
    char home[4096];
    if (!getcwd(home, sizeof home)) strcpy(home, "/");
    for (;;) {
        const int fd = accept4(ls, nullptr, nullptr, SOCK_CLOEXEC);
        if (fd < 0) { if (errno == EINTR) continue; break; }
        uint32_t len = 0;
        msghdr m;
        memset(&m, 0, sizeof m);
        iovec io = {&len, 4};
        m.msg_iov = &io;
        m.msg_iovlen = 1;
        char ctl[CMSG_SPACE(2 * sizeof(int))];
        m.msg_control = ctl;
        m.msg_controllen = sizeof ctl;
        int cfd[2] = {-1, -1};
        const ssize_t got = recvmsg(fd, &m, MSG_CMSG_CLOEXEC);
        for (cmsghdr* c = CMSG_FIRSTHDR(&m); got == 4 && c; c = CMSG_NXTHDR(&m, c))
            if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS && c->cmsg_len >= CMSG_LEN(2 * sizeof(int)))
                memcpy(cfd, CMSG_DATA(c), sizeof cfd);
        std::string body;
        bool ok = got == 4 && cfd[0] >= 0 && cfd[1] >= 0 && len >= 16 && len < (64u << 20);
        if (ok) { body.resize(len); ok = read_all(fd, &body[0], len); }
        uint32_t head[4] = {0, 0, 0, 0};
        std::vector<char*> strs;
        if (ok) {
            memcpy(head, body.data(), 16);
            ok = head[0] == kMagic && head[1] <= 2 && body.back() == '\0';
            for (size_t o = 16; ok && o < body.size(); o += strlen(&body[o]) + 1) strs.push_back(&body[o]);
            ok = ok && strs.size() == 1 + (size_t)head[2] + (size_t)head[3] && head[2] >= 1;
        }
        if (!ok) {
            if (cfd[0] >= 0) close(cfd[0]);
            if (cfd[1] >= 0) close(cfd[1]);
            close(fd);
            continue;
        }
        if (!write_all(fd, &kMagic, 4)) { close(cfd[0]); close(cfd[1]); close(fd); continue; }
        // the program runs here, writing to the client's descriptors, in the client's directory and AS_* environment
        std::cout.flush();
        std::cerr.flush();
        fflush(nullptr);
        const int save1 = dup(1), save2 = dup(2);
        dup2(cfd[0], 1);
        dup2(cfd[1], 2);
        close(cfd[0]);
        close(cfd[1]);
        std::vector<std::string> set_names;
        for (uint32_t i = 0; i < head[3]; ++i) {
            char* kv = strs[1 + head[2] + i];
            if (char* eq = strchr(kv, '=')) {
                std::string name(kv, eq);
                setenv(name.c_str(), eq + 1, 1);
                set_names.push_back(name);
            }
        }
        int32_t rc = 1;
        if (chdir(strs[0]) == 0) {
            std::vector<char*> av(strs.begin() + 1, strs.begin() + 1 + head[2]);
            av.push_back(nullptr);
            rc = run_program(head[1], (int)head[2], av.data());
        } else {
            fprintf(stderr, "amplisolve_b200_serve: cannot enter %s\n", strs[0]);
        }
        std::cout.flush();
        std::cerr.flush();
        fflush(nullptr);
        for (const std::string& n : set_names) unsetenv(n.c_str());
        if (chdir(home) != 0) {}
        dup2(save1, 1);
        dup2(save2, 2);
        close(save1);
        close(save2);
        write_all(fd, &rc, 4);
        close(fd);
    }
    return 0;
}
