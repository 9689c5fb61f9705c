// Test helper (no GPU needed) for the multi-GPU plumbing of the host layer:
//   shard_check shards WORLD off0 off1 ... offQ     prints "q_begin q_end" per rank (host::query_shards)
//   shard_check rendezvous WORLD PORT [ADDR]        forks WORLD processes; rank 0 hands a 128-byte token to the
//                                                   others over TCP (host::exchange_bytes); exit 0 iff all got it
#include <sys/wait.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "quickrank_host.h"

using namespace quickrank;

int main(int argc, char **argv) {
  if (argc >= 4 && !strcmp(argv[1], "shards")) {
    const int world = atoi(argv[2]);
    std::vector<uint64_t> off;
    for (int i = 3; i < argc; ++i) off.push_back(strtoull(argv[i], nullptr, 10));
    for (auto &s : host::query_shards(off.data(), off.size() - 1, world)) printf("%zu %zu\n", s.first, s.second);
    return 0;
  }
  if ((argc == 4 || argc == 5) && !strcmp(argv[1], "rendezvous")) {
    const int world = atoi(argv[2]), port = atoi(argv[3]);
    const std::string addr = argc == 5 ? argv[4] : "127.0.0.1";
    int rank = 0;
    std::vector<pid_t> kids;
    for (int r = 1; r < world; ++r) {
      const pid_t pid = fork();
      if (pid == 0) { rank = r; kids.clear(); break; }
      kids.push_back(pid);
    }
    host::Sharding s;
    s.rank = rank; s.world = world; s.local_rank = rank; s.port = port; s.addr = addr;
    unsigned char id[128];
    for (int i = 0; i < 128; ++i) id[i] = rank == 0 ? (unsigned char) (i * 7 + 3) : 0;
    bool ok = host::exchange_bytes(id, sizeof(id), s, 20);
    for (int i = 0; i < 128 && ok; ++i) ok = id[i] == (unsigned char) (i * 7 + 3);
    if (rank != 0) return ok ? 0 : 1;
    for (pid_t k : kids) {
      int st = 0;
      waitpid(k, &st, 0);
      if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) ok = false;
    }
    puts(ok ? "ok" : "FAILED");
    return ok ? 0 : 1;
  }
  std::cerr << "usage: shard_check shards WORLD off0 ... offQ | shard_check rendezvous WORLD PORT [ADDR]" << std::endl;
  return 2;
}
