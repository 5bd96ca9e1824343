// owc_bench.cc -- BASELINE config 1: OrderedWordCount (tez-examples/src/main/java/org/apache/tez/examples/
// OrderedWordCount.java:124-180) on ~100 MB of synthetic text with 4 reducers, driven through the plugin mirror.
//
//   tokenizer tasks  --(Text word, IntWritable 1), HashPartitioner, OrderedPartitionedKVOutput-->  summation tasks
//   summation tasks  --(IntWritable count, Text word), 1 partition, OrderedPartitionedKVOutput-->   sorter task
//
//   owc_bench gpu <text_mb> <tokenizers> <reducers> <workdir>   both edges through libtezgpu's plugin mirror
//                                                               (include/tez_runtime.h: files, indexes, counters as Tez)
//   owc_bench cpu <text_mb> <tokenizers> <reducers> <workdir>   the same job through the CPU restatement (oracle/): the
//                                                               reference's PipelinedSorter + TezMerger arm, one thread
//                                                               per task phase like local mode
// Both arms check the job's known answer (the word histogram of the generated text, ascending by count) and print one
// JSON object with the wall time of the shuffle-bound part (sorting, spilling, merging, grouping; text generation and
// tokenisation excluded).  Test / benchmark infrastructure: the only program that links both libraries.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "tez_runtime.h"
#include "tezgpu.h"
extern "C" {
#include "tez_oracle.h"
}

static double now() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}
static void die(const char *what, const char *msg) {
  fprintf(stderr, "owc_bench: %s: %s\n", what, msg ? msg : "");
  exit(1);
}
#define RT(x) do { if ((x) != 0) die(#x, tezrt_last_error()); } while (0)

struct Job {
  int tokenizers, reducers;
  uint32_t vocab;
  std::vector<std::vector<uint32_t>> tokens;  // word ids per tokenizer task
  std::vector<uint64_t> histogram;            // expected count per word id
  uint64_t text_bytes = 0, kv_bytes = 0, records = 0;
};

static std::string word_of(uint32_t id) {
  char b[16];
  snprintf(b, sizeof(b), "w%05u", id);
  return b;
}
static std::string text_key(const std::string &w) { return std::string(1, (char)w.size()) + w; }  // Text: vint(len) + UTF-8
static std::string int_be(uint32_t v) {
  char b[4] = {(char)(v >> 24), (char)(v >> 16), (char)(v >> 8), (char)v};
  return std::string(b, 4);
}

// Zipf(1.0) over `vocab` words, counter-based: token i = f(seed, i)
static void generate(Job &j, uint64_t text_mb, uint64_t seed) {
  j.vocab = 1000;
  std::vector<double> cdf(j.vocab);
  double s = 0;
  for (uint32_t k = 0; k < j.vocab; k++) { s += 1.0 / (double)(k + 1); cdf[k] = s; }
  for (auto &c : cdf) c /= s;
  const uint64_t ntok = (text_mb << 20) / 7;  // "w00000 " = 7 bytes of text per token
  j.tokens.assign((size_t)j.tokenizers, {});
  j.histogram.assign(j.vocab, 0);
  for (uint64_t i = 0; i < ntok; i++) {
    const double u = (double)(tzo_splitmix64((seed << 56) ^ i) >> 11) * (1.0 / 9007199254740992.0);
    const uint32_t id = (uint32_t)(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin());
    j.tokens[(size_t)(i * (uint64_t)j.tokenizers / ntok)].push_back(id < j.vocab ? id : j.vocab - 1);
    j.histogram[id < j.vocab ? id : j.vocab - 1]++;
  }
  j.text_bytes = ntok * 7;
  j.records = ntok;
  j.kv_bytes = ntok * (7 + 4);  // Text key 1+6 bytes, IntWritable value 4 bytes
}

static void check_answer(const Job &j, const std::vector<std::pair<uint32_t, std::string>> &final_rows) {
  std::vector<std::pair<uint32_t, std::string>> want;
  for (uint32_t id = 0; id < j.vocab; id++)
    if (j.histogram[id]) want.push_back({(uint32_t)j.histogram[id], word_of(id)});
  std::sort(want.begin(), want.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
  if (final_rows.size() != want.size()) die("answer", "wrong number of rows");
  for (size_t i = 0; i < want.size(); i++)
    if (final_rows[i].first != want[i].first) die("answer", "counts not in ascending order / wrong count");
  std::vector<std::pair<uint32_t, std::string>> a = final_rows, b = want;   // words with equal counts come in any order
  std::sort(a.begin(), a.end());
  std::sort(b.begin(), b.end());
  if (a != b) die("answer", "word/count pairs differ from the histogram of the text");
}

// ------------------------------------------------------------------------------------------------------------ GPU arm
static double run_gpu(const Job &j, const std::string &work) {
  const std::string conf1 =
      "tez.runtime.key.class=org.apache.hadoop.io.Text\ntez.runtime.value.class=org.apache.hadoop.io.IntWritable\n"
      "tez.runtime.io.sort.mb=1024\n";
  const std::string conf2 =
      "tez.runtime.key.class=org.apache.hadoop.io.IntWritable\ntez.runtime.value.class=org.apache.hadoop.io.Text\n"
      "tez.runtime.io.sort.mb=64\n";
  const int64_t task_mem = 8ll << 30;
  const double t0 = now();
  std::vector<tezrt_output *> toks((size_t)j.tokenizers);
  const std::string one = int_be(1);
  for (int t = 0; t < j.tokenizers; t++) {
    const std::string wd = work + "/tok" + std::to_string(t);
    mkdir(wd.c_str(), 0755);
    RT(tezrt_output_create(conf1.c_str(), wd.c_str(), ("attempt_tok_" + std::to_string(t)).c_str(), "summation", "localhost", 13562,
                           task_mem, j.reducers, 0, &toks[t]));
    int64_t req = 0;
    RT(tezrt_output_initialize(toks[t], &req));
    RT(tezrt_output_memory_assigned(toks[t], req));
    RT(tezrt_output_start(toks[t]));
    std::vector<std::string> keys(j.vocab);
    for (uint32_t id = 0; id < j.vocab; id++) keys[id] = text_key(word_of(id));
    for (uint32_t id : j.tokens[t])
      RT(tezrt_output_write(toks[t], (const uint8_t *)keys[id].data(), (uint32_t)keys[id].size(), (const uint8_t *)one.data(), 4, -1));
    int32_t nev = 0;
    RT(tezrt_output_close(toks[t], &nev));
  }
  std::vector<tezrt_output *> sums((size_t)j.reducers);
  for (int r = 0; r < j.reducers; r++) {
    const std::string wd = work + "/sum" + std::to_string(r);
    mkdir(wd.c_str(), 0755);
    tezrt_input *in = nullptr;
    RT(tezrt_input_create(conf1.c_str(), wd.c_str(), ("attempt_sum_" + std::to_string(r)).c_str(), task_mem, j.tokenizers, 0, &in));
    int64_t req = 0;
    RT(tezrt_input_initialize(in, &req));
    RT(tezrt_input_start(in));
    for (int t = 0; t < j.tokenizers; t++)
      RT(tezrt_input_add_local_output(in, t, tezrt_output_file(toks[t]), tezrt_output_index_file(toks[t]), r, 0));
    RT(tezrt_input_wait_ready(in));
    RT(tezrt_output_create(conf2.c_str(), wd.c_str(), ("attempt_sumout_" + std::to_string(r)).c_str(), "sorter", "localhost", 13562, task_mem,
                           1, 0, &sums[r]));
    RT(tezrt_output_initialize(sums[r], &req));
    RT(tezrt_output_memory_assigned(sums[r], req));
    RT(tezrt_output_start(sums[r]));
    const uint8_t *k, *v;
    uint32_t kl, vl;
    int rc;
    while ((rc = tezrt_input_next(in, &k, &kl)) == 1) {
      const std::string word((const char *)k, kl);
      uint32_t sum = 0;
      while ((rc = tezrt_input_next_value(in, &v, &vl)) == 1) sum += ((uint32_t)v[0] << 24) | ((uint32_t)v[1] << 16) | ((uint32_t)v[2] << 8) | v[3];
      if (rc < 0) die("next_value", tezrt_last_error());
      const std::string cnt = int_be(sum);
      RT(tezrt_output_write(sums[r], (const uint8_t *)cnt.data(), 4, (const uint8_t *)word.data(), (uint32_t)word.size(), -1));
    }
    if (rc < 0) die("next", tezrt_last_error());
    int32_t nev = 0;
    RT(tezrt_output_close(sums[r], &nev));
    tezrt_input_destroy(in);
  }
  std::vector<std::pair<uint32_t, std::string>> rows;
  {
    const std::string wd = work + "/sorter";
    mkdir(wd.c_str(), 0755);
    tezrt_input *in = nullptr;
    RT(tezrt_input_create(conf2.c_str(), wd.c_str(), "attempt_sorter_0", task_mem, j.reducers, 0, &in));
    int64_t req = 0;
    RT(tezrt_input_initialize(in, &req));
    RT(tezrt_input_start(in));
    for (int r = 0; r < j.reducers; r++)
      RT(tezrt_input_add_local_output(in, r, tezrt_output_file(sums[r]), tezrt_output_index_file(sums[r]), 0, 0));
    RT(tezrt_input_wait_ready(in));
    const uint8_t *k, *v;
    uint32_t kl, vl;
    int rc;
    while ((rc = tezrt_input_next(in, &k, &kl)) == 1) {
      const uint32_t cnt = ((uint32_t)k[0] << 24) | ((uint32_t)k[1] << 16) | ((uint32_t)k[2] << 8) | k[3];
      while ((rc = tezrt_input_next_value(in, &v, &vl)) == 1) rows.push_back({cnt, std::string((const char *)v + 1, vl - 1)});
    }
    tezrt_input_destroy(in);
  }
  const double secs = now() - t0;
  for (auto *o : toks) tezrt_output_destroy(o);
  for (auto *o : sums) tezrt_output_destroy(o);
  check_answer(j, rows);
  return secs;
}

// ------------------------------------------------------------------------------------------------------------ CPU arm
struct Packed {
  std::vector<uint8_t> kv;
  std::vector<uint64_t> ko;
  std::vector<uint32_t> kl, vl;
  void add(const std::string &k, const std::string &v) {
    ko.push_back(kv.size());
    kv.insert(kv.end(), k.begin(), k.end());
    kv.insert(kv.end(), v.begin(), v.end());
    kl.push_back((uint32_t)k.size());
    vl.push_back((uint32_t)v.size());
  }
};

static double run_cpu(const Job &j) {
  const double t0 = now();
  const std::string one = int_be(1);
  std::vector<tzo_sorter_result> tok((size_t)j.tokenizers);
  std::vector<std::string> keys(j.vocab);
  for (uint32_t id = 0; id < j.vocab; id++) keys[id] = text_key(word_of(id));
  for (int t = 0; t < j.tokenizers; t++) {
    Packed p;
    for (uint32_t id : j.tokens[t]) p.add(keys[id], one);
    tzo_sorter_conf c;
    memset(&c, 0, sizeof(c));
    c.num_partitions = j.reducers; c.cmp_kind = TZO_CMP_TEXT; c.partitioner = TZO_PART_HASH; c.send_empty_partition_details = 1;
    c.rle_policy = -1; c.sort_threads = 2;   // tez.runtime.pipelined.sorter.sort.threads default
    if (tzo_pipelined_sort(&c, p.kv.data(), p.ko.data(), p.kl.data(), p.vl.data(), nullptr, p.ko.size(), &tok[t]) != 0) die("tzo_pipelined_sort", "");
  }
  std::vector<tzo_sorter_result> sums((size_t)j.reducers);
  for (int r = 0; r < j.reducers; r++) {
    std::vector<tzo_segment> segs;
    for (int t = 0; t < j.tokenizers; t++) {
      const int64_t start = tok[t].index[3 * r], raw = tok[t].index[3 * r + 1], part = tok[t].index[3 * r + 2];
      if (raw > 6) segs.push_back({tok[t].file_out.data + start, (size_t)part, 1});
    }
    tzo_merge_result m;
    if (tzo_merge(segs.data(), (int)segs.size(), TZO_CMP_TEXT, 100, 0, 1, 0, &m) != 0) die("tzo_merge", "");
    Packed p;
    uint64_t koff = 0, voff = 0;
    std::string cur;
    uint32_t sum = 0;
    bool have = false;
    for (uint64_t i = 0; i < m.n; i++) {
      const std::string k((const char *)m.keys.data + koff, m.key_len[i]);
      const uint8_t *v = m.vals.data + voff;
      koff += m.key_len[i];
      voff += m.val_len[i];
      if (!have || (!m.same_key[i] && k != cur)) {   // ValuesIterator.readNextKey (RL/common/ValuesIterator.java:177-201)
        if (have) p.add(int_be(sum), cur);
        cur = k; sum = 0; have = true;
      }
      sum += ((uint32_t)v[0] << 24) | ((uint32_t)v[1] << 16) | ((uint32_t)v[2] << 8) | v[3];
    }
    if (have) p.add(int_be(sum), cur);
    tzo_merge_result_free(&m);
    tzo_sorter_conf c;
    memset(&c, 0, sizeof(c));
    c.num_partitions = 1; c.cmp_kind = TZO_CMP_INT; c.partitioner = TZO_PART_HASH; c.send_empty_partition_details = 1; c.rle_policy = -1;
    c.sort_threads = 2;
    if (tzo_pipelined_sort(&c, p.kv.data(), p.ko.data(), p.kl.data(), p.vl.data(), nullptr, p.ko.size(), &sums[r]) != 0) die("tzo_pipelined_sort", "");
  }
  std::vector<std::pair<uint32_t, std::string>> rows;
  {
    std::vector<tzo_segment> segs;
    for (int r = 0; r < j.reducers; r++)
      if (sums[r].index[1] > 6) segs.push_back({sums[r].file_out.data + sums[r].index[0], (size_t)sums[r].index[2], 1});
    tzo_merge_result m;
    if (tzo_merge(segs.data(), (int)segs.size(), TZO_CMP_INT, 100, 0, 1, 0, &m) != 0) die("tzo_merge", "");
    uint64_t koff = 0, voff = 0;
    for (uint64_t i = 0; i < m.n; i++) {
      const uint8_t *k = m.keys.data + koff;
      rows.push_back({((uint32_t)k[0] << 24) | ((uint32_t)k[1] << 16) | ((uint32_t)k[2] << 8) | k[3],
                      std::string((const char *)m.vals.data + voff + 1, m.val_len[i] - 1)});
      koff += m.key_len[i];
      voff += m.val_len[i];
    }
    tzo_merge_result_free(&m);
  }
  const double secs = now() - t0;
  for (auto &r : tok) tzo_sorter_result_free(&r);
  for (auto &r : sums) tzo_sorter_result_free(&r);
  check_answer(j, rows);
  return secs;
}

int main(int argc, char **argv) {
  if (argc < 6) die("usage", "owc_bench gpu|cpu <text_mb> <tokenizers> <reducers> <workdir>");
  Job j;
  const std::string mode = argv[1];
  const uint64_t mb = (uint64_t)atoll(argv[2]);
  j.tokenizers = atoi(argv[3]);
  j.reducers = atoi(argv[4]);
  generate(j, mb, 1);
  double secs;
  if (mode == "gpu") {
    secs = run_gpu(j, argv[5]);      // warm-up pass (allocations, first-touch), then the timed one
    const std::string w2 = std::string(argv[5]) + "/timed";
    mkdir(w2.c_str(), 0755);
    secs = run_gpu(j, w2);
  } else {
    secs = run_cpu(j);
  }
  printf("{\"arm\": \"%s\", \"text_bytes\": %llu, \"records\": %llu, \"kv_bytes\": %llu, \"tokenizers\": %d, \"reducers\": %d, "
         "\"seconds\": %.4f, \"kv_gbs\": %.5f, \"answer_checked\": true}\n",
         mode.c_str(), (unsigned long long)j.text_bytes, (unsigned long long)j.records, (unsigned long long)j.kv_bytes, j.tokenizers,
         j.reducers, secs, (double)j.kv_bytes / secs / 1e9);
  return 0;
}
