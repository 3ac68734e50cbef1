// Experiment (GPU box, N GPUs): what does the HOST side of the end-to-end path deliver when several ranks copy their
// input batch (154 MB of pinned fp32) to their own GPU at the same time?  bench.py's e2e number at 8 GPUs is bound by
// exactly this copy.  One process per GPU (fork before any CUDA call, like torchrun's ranks); a
// process-shared barrier lines the ranks up, then each times ITERS back-to-back cudaMemcpyAsync of BYTES with CUDA
// events.  Reported per configuration: per-rank GB/s (min / median / max) and the sum.
//   sets:    {0} {0,1} {0..3} {0..7}, then pairs {0,k} (which GPUs share a PCIe uplink?), then {4..7}
//   allocs:  cudaHostAlloc default | write-combined | after binding the process to the GPU's local CPUs (NUMA-local pages)
//   nvcc -O2 -o h2d_concurrent h2d_concurrent.cu && ./h2d_concurrent [bytes] [iters]
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct Shared {
  pthread_barrier_t bar;
  double gbs[16];
  int ok[16];
};

static std::string read_line(const std::string& path) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return "?";
  char buf[512] = "";
  if (!fgets(buf, sizeof(buf), f)) buf[0] = 0;
  fclose(f);
  std::string s(buf);
  while (!s.empty() && (s.back() == '\n' || s.back() == ' ')) s.pop_back();
  return s;
}

static void bind_to_cpulist(const std::string& list) {   // "0-15,32-47"
  cpu_set_t set;
  CPU_ZERO(&set);
  const char* p = list.c_str();
  bool any = false;
  while (*p) {
    char* e;
    long a = strtol(p, &e, 10);
    if (e == p) break;
    long b = a;
    if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
    for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET(c, &set); any = true; }
    p = (*e == ',') ? e + 1 : e;
    if (*e != ',') break;
  }
  if (any) sched_setaffinity(0, sizeof(set), &set);
}

enum Alloc { DEFAULT = 0, WRITE_COMBINED = 1, NUMA_LOCAL = 2 };

static void rank_main(Shared* sh, int slot, int dev, size_t bytes, int iters, int alloc) {
  sh->ok[slot] = 0;
  if (cudaSetDevice(dev) != cudaSuccess) { pthread_barrier_wait(&sh->bar); pthread_barrier_wait(&sh->bar); _exit(1); }
  if (alloc == NUMA_LOCAL) {
    char bus[64] = "";
    cudaDeviceGetPCIBusId(bus, sizeof(bus), dev);
    for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
    bind_to_cpulist(read_line(std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist"));
  }
  float* h = nullptr;
  float* d = nullptr;
  cudaHostAlloc((void**)&h, bytes, alloc == WRITE_COMBINED ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
  cudaMalloc((void**)&d, bytes);
  if (h) memset(h, 1, bytes);
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
  pthread_barrier_wait(&sh->bar);
  cudaEventRecord(e0, st);
  for (int i = 0; i < iters; ++i) cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
  cudaEventRecord(e1, st);
  cudaError_t e = cudaStreamSynchronize(st);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  sh->gbs[slot] = (e == cudaSuccess && ms > 0) ? (double)bytes * iters / (ms * 1e-3) / 1e9 : 0.0;
  sh->ok[slot] = e == cudaSuccess;
  pthread_barrier_wait(&sh->bar);
  _exit(0);
}

static void run_set(const std::vector<int>& devs, size_t bytes, int iters, int alloc) {
  Shared* sh = (Shared*)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  pthread_barrierattr_t at;
  pthread_barrierattr_init(&at);
  pthread_barrierattr_setpshared(&at, PTHREAD_PROCESS_SHARED);
  pthread_barrier_init(&sh->bar, &at, (unsigned)devs.size());
  std::vector<pid_t> pids;
  for (size_t i = 0; i < devs.size(); ++i) {
    pid_t p = fork();
    if (p == 0) rank_main(sh, (int)i, devs[i], bytes, iters, alloc);
    pids.push_back(p);
  }
  for (pid_t p : pids) { int st; waitpid(p, &st, 0); }
  std::vector<double> g(sh->gbs, sh->gbs + devs.size());
  std::string names;
  for (int d : devs) names += std::to_string(d) + " ";
  double sum = 0;
  for (double v : g) sum += v;
  std::vector<double> s = g;
  std::sort(s.begin(), s.end());
  static const char* an[] = {"default", "write-combined", "numa-local"};
  printf("gpus [%s] alloc=%-14s per-rank GB/s min %.1f med %.1f max %.1f  sum %.1f   (", names.c_str(), an[alloc], s.front(), s[s.size() / 2], s.back(), sum);
  for (size_t i = 0; i < g.size(); ++i) printf("%s%.1f", i ? " " : "", g[i]);
  printf(")\n");
  fflush(stdout);
  munmap(sh, sizeof(Shared));
}

int main(int argc, char** argv) {
  const size_t bytes = argc > 1 ? (size_t)atoll(argv[1]) : (size_t)256 * 3 * 224 * 224 * 4;
  const int iters = argc > 2 ? atoi(argv[2]) : 20;
  // device count WITHOUT creating a CUDA context in the parent (fork + CUDA do not mix): ask nvidia-smi
  int n = 0;
  if (FILE* f = popen("nvidia-smi -L | wc -l", "r")) { if (fscanf(f, "%d", &n) != 1) n = 0; pclose(f); }
  printf("h2d_concurrent: %d GPUs, %.1f MB per copy, %d copies per rank per measurement\n", n, bytes / 1e6, iters);
  fflush(stdout);
  if (system("nvidia-smi topo -m 2>&1 | head -30; echo; lscpu | grep -i -E 'numa|model name|^cpu\\(s\\)|socket'; echo; numactl --hardware 2>&1 | head -12; echo; "
             "for d in /sys/bus/pci/devices/*; do if [ \"$(cat $d/vendor 2>/dev/null)\" = 0x10de ] && [ \"$(cat $d/class 2>/dev/null | cut -c1-6)\" = 0x0302 ]; "
             "then echo \"$d numa_node=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)\"; fi; done; "
             "echo; grep -E 'MemTotal|MemFree' /proc/meminfo; nproc") != 0) {}
  fflush(stdout);
  if (n < 1) return 1;
  std::vector<std::vector<int>> sets;
  sets.push_back({0});
  for (int k : {2, 4, 8}) if (n >= k) { std::vector<int> s; for (int i = 0; i < k; ++i) s.push_back(i); sets.push_back(s); }
  for (auto& s : sets)
    for (int alloc : {DEFAULT, WRITE_COMBINED, NUMA_LOCAL}) run_set(s, bytes, iters, alloc);
  for (int k = 1; k < n; ++k) run_set({0, k}, bytes, iters, DEFAULT);
  if (n >= 8) { run_set({4, 5, 6, 7}, bytes, iters, DEFAULT); run_set({0, 2, 4, 6}, bytes, iters, DEFAULT); run_set({1, 3, 5, 7}, bytes, iters, DEFAULT); }
  return 0;
}
