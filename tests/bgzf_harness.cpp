// CPU harness (built and run by tests/test_host_cpu.py): the front end's whole-stream BGZF writer -- blocks deflated in parallel --
// against its sequential write()/flush()/tell() path, which the BAM golden files pin: same bytes, same virtual addresses.
#define main bkx_align_main_entry
#include "../biokanga_b200/csrc/host/bkx_align_main.cpp"
#undef main
#include <random>
int main() {
  std::mt19937_64 rng(7);
  int fails = 0;
  for (int trial = 0; trial < 60; ++trial) {
    // random record stream
    size_t nrec = 1 + rng() % 3000;
    std::vector<std::vector<uint8_t>> recs(nrec);
    std::vector<uint8_t> U;
    std::vector<uint64_t> ofs;
    size_t hdr = rng() % 5000;
    U.resize(hdr);
    for (auto& b : U) b = (uint8_t)(rng() % 7);
    for (auto& r : recs) {
      r.resize(30 + rng() % 400);
      for (auto& b : r) b = (uint8_t)("ACGT"[rng() % 4]);
      ofs.push_back(U.size());
      U.insert(U.end(), r.begin(), r.end());
    }
    bool with_flush = trial % 3 != 0;
    size_t fl_rec = rng() % nrec;                        // explicit flush behind this record
    if (trial % 7 == 1) fl_rec = nrec - 1;               // ... sometimes the very last one
    uint64_t flush_at = ofs[fl_rec] + recs[fl_rec].size();
    // sequential reference behaviour
    Bgzf a;
    a.open("bgzf_a.bin");
    a.write(U.data(), hdr);
    std::vector<uint64_t> sva(nrec), eva(nrec);
    for (size_t i = 0; i < nrec; ++i) {
      sva[i] = a.tell();
      a.write(recs[i].data(), recs[i].size());
      if (with_flush && i == fl_rec) a.flush();
      eva[i] = a.tell();
    }
    a.close();
    Bgzf b;
    b.open("bgzf_b.bin");
    b.write_stream(U.data(), U.size(), with_flush ? &flush_at : nullptr, 1 + trial % 8);
    bool ok = true;
    for (size_t i = 0; i < nrec; ++i)
      if (b.vaddr(ofs[i]) != sva[i] || b.vaddr(ofs[i] + recs[i].size()) != eva[i]) { ok = false; break; }
    b.close();
    auto slurpf = [](const char* p) { std::vector<char> v; FILE* f = fopen(p, "rb"); char buf[65536]; size_t n; while ((n = fread(buf, 1, sizeof buf, f)) > 0) v.insert(v.end(), buf, buf + n); fclose(f); return v; };
    if (!ok || slurpf("bgzf_a.bin") != slurpf("bgzf_b.bin")) { ++fails; printf("trial %d MISMATCH (va ok %d)\n", trial, (int)ok); }
  }
  printf("%s\n", fails ? "FAIL" : "all 60 trials identical");
  return fails != 0;
}
