// Host-side assembly of decoded peak records into V2000 MOL-block text: native,
// multi-threaded restatement of the per-image Python loop of the reference, so that the stage after the decode kernel
// keeps up with the GPU. No device code in this file.
//
// Follows, statement by statement,
//   /root/reference/src/img2smiles.py:183-187  greedy < 2 px de-duplication of atom peaks
//   /root/reference/src/img2smiles.py:195-212  bond end points -> nearest atoms by the anisotropic distance (float64 numpy)
//   /root/reference/src/img2smiles.py:214-236  pair de-duplication            :249-274  valence repair
//   /root/reference/src/img2smiles.py:276-314  re-indexing, implicit-H list of aromatic hetero atoms
//   /root/reference/src/generate_smiles.py:18-105  MOL-block text (RDKit parses this text at :115-118)
// The float64 arithmetic reproduces numpy's operation order exactly (no FMA contraction: see the Makefile flag), the
// cos / sin of the 60 omega bins come from the caller (computed by numpy, as abcnet_b200.records_to_lists does), and
// np.argmin's first-minimum / NaN rule is restated, so the text is byte-identical to the Python path on the same records
// (tests/test_assemble_cpu.py pins it against MOL blocks minted by the reference's own code, tests/golden).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace abc {
namespace {

// utils.py:12-14 inverted as in img2smiles.py:24-26 (index 0 -> 'C'), followed by the symbols only valence repair produces
const char* const kSymbols[] = {"C", "C", "N", "O", "P", "F", "Cl", "S", "Br", "B", "Se", "I", "H", "Si"};
constexpr int kNumSymbols = 14;
const int kChargeValues[3] = {0, 1, -1};                      // img2smiles.py:26

int max_valence(const char* s) {                              // img2smiles.py:30-32 (atom_max_valence)
  struct E { const char* s; int v; };
  static const E tab[] = {{"O", 2}, {"C", 4}, {"N", 3}, {"F", 1}, {"H", 1}, {"S", 6}, {"Cl", 1}, {"P", 5}, {"Br", 1}, {"B", 3},
                          {"I", 1}, {"Si", 4}, {"Se", 6}, {"Te", 6}, {"As", 3}, {"Al", 3}, {"Zn", 2}, {"Ca", 2}, {"Ag", 1}};
  for (const E& e : tab)
    if (!std::strcmp(e.s, s)) return e.v;
  return 4;
}

// np.argmin over a row: first minimum; a NaN wins and stops the scan (numpy DOUBLE_argmin)
int np_argmin(const std::vector<double>& v) {
  double mp = v[0];
  int idx = 0;
  if (std::isnan(mp)) return 0;
  for (size_t i = 1; i < v.size(); ++i) {
    if (!(v[i] >= mp)) {
      mp = v[i];
      idx = static_cast<int>(i);
      if (std::isnan(mp)) break;
    }
  }
  return idx;
}

inline double lrelu_half(double v) { return std::fmax(v, 0.5 * v); }      // img2smiles.py leaky_relu (slope 0.5); NaN-safe like np.maximum? see below

void append_rjust(std::string& t, const std::string& s, size_t w) {
  if (s.size() < w) t.append(w - s.size(), ' ');
  t += s;
}

void append_coord(std::string& t, double p) {                 // generate_smiles.py:31-38
  char buf[64];
  if (p < 0) {
    std::snprintf(buf, sizeof(buf), "   %2.4f", p);
  } else {
    std::snprintf(buf, sizeof(buf), "    %.4f", p);
  }
  t += buf;
}

// One image. Returns false for "no molecule" (img2smiles.py:126-129 / empty lists).
bool assemble_one(const AbcAtomRec* atoms, int na, const AbcBondRec* bonds, int nb, int n_bond_peaks, const double* cos_tab,
                  const double* sin_tab, int n_omega, std::string& text) {
  if (na == 0 || n_bond_peaks == 0) return false;
  // ---- atom peaks, greedy de-duplication in enumeration order (img2smiles.py:183-187)
  std::vector<int> ax, ay, ahs, acharge;
  std::vector<const char*> atype;
  for (int i = 0; i < na; ++i) {
    const int x = atoms[i].x, y = atoms[i].y;
    bool dup = false;
    for (size_t k = 0; k < ax.size() && !dup; ++k) {
      const long dx = ax[k] - x, dy = ay[k] - y;
      dup = dx * dx + dy * dy < 4;
    }
    if (dup) continue;
    ax.push_back(x);
    ay.push_back(y);
    atype.push_back(kSymbols[atoms[i].type < kNumSymbols ? atoms[i].type : 0]);
    acharge.push_back(kChargeValues[atoms[i].charge < 3 ? atoms[i].charge : 0]);
    ahs.push_back(atoms[i].hs);
  }
  const int n_at = static_cast<int>(ax.size());
  if (nb == 0 || n_at == 0) return false;
  // ---- bond -> atom assignment (img2smiles.py:195-212)
  std::vector<std::pair<int, int>> pairs;
  std::vector<int> orders;
  std::vector<double> d_a(n_at), d_b(n_at);
  for (int i = 0; i < nb; ++i) {
    const int w = bonds[i].omega < n_omega ? bonds[i].omega : 0;
    const double rho = static_cast<double>(bonds[i].rho);
    const double bdx = rho * cos_tab[w], bdy = rho * sin_tab[w];
    const double bpx = static_cast<double>(bonds[i].x), bpy = static_cast<double>(bonds[i].y);
    const double eax = bpx + bdx, eay = bpy + bdy, ebx = bpx - bdx, eby = bpy - bdy;
    const double nrm = std::sqrt(bdx * bdx + bdy * bdy);
    const double ux = bdx / nrm, uy = bdy / nrm;
    const double vx = -uy, vy = ux;                                     // flip, then negate the first component
    for (int j = 0; j < n_at; ++j) {
      const double axj = static_cast<double>(ax[j]), ayj = static_cast<double>(ay[j]);
      const double pax = eax - axj, pay = eay - ayj, pbx = ebx - axj, pby = eby - ayj;
      const double sa = pax * ux + pay * uy, sb = -(pbx * ux + pby * uy);
      const double ta = (2.0 * pax) * vx + (2.0 * pay) * vy, tb = (2.0 * pbx) * vx + (2.0 * pby) * vy;
      // np.maximum propagates NaN; std::fmax would drop it -> handle explicitly
      const double la = std::isnan(sa) ? sa : lrelu_half(sa), lb = std::isnan(sb) ? sb : lrelu_half(sb);
      d_a[j] = std::fabs(la) + std::fabs(ta);
      d_b[j] = std::fabs(lb) + std::fabs(tb);
    }
    const int a = np_argmin(d_b);                                       // img2smiles.py:211 (sic: index1 from distance2)
    const int b = np_argmin(d_a);                                       // img2smiles.py:212
    if (a == b) continue;
    bool seen = false;
    for (const auto& pr : pairs)
      if ((pr.first == a && pr.second == b) || (pr.first == b && pr.second == a)) {
        seen = true;
        break;
      }
    if (seen) continue;
    pairs.emplace_back(a, b);
    orders.push_back(static_cast<int>(bonds[i].type) + 1);              // bond_type_devocab, img2smiles.py:28
  }
  // ---- valence repair (img2smiles.py:249-274)
  std::vector<char> used(n_at, 0);
  std::vector<int> load(n_at);
  for (int i = 0; i < n_at; ++i) load[i] = -acharge[i];
  for (size_t k = 0; k < pairs.size(); ++k) {
    used[pairs[k].first] = used[pairs[k].second] = 1;
    const int o = orders[k];
    const int n = (o == 4 || o == 5 || o == 6) ? 1 : o;
    load[pairs[k].first] += n;
    load[pairs[k].second] += n;
  }
  static const char* const repair[8] = {nullptr, nullptr, "O", "N", "C", "P", "S", "Cl"};
  for (int i = 0; i < n_at; ++i)
    if (max_valence(atype[i]) < load[i] && load[i] >= 2 && load[i] <= 7) atype[i] = repair[load[i]];
  // ---- re-indexing (img2smiles.py:276-300)
  std::vector<int> remap(n_at), f_x, f_y, f_charge, f_hs;
  std::vector<const char*> f_type;
  int k1 = 1;
  for (int i = 0; i < n_at; ++i) {
    remap[i] = k1;
    if (used[i]) {
      f_type.push_back(atype[i]);
      f_charge.push_back(acharge[i]);
      f_x.push_back(ax[i]);
      f_y.push_back(ay[i]);
      f_hs.push_back(ahs[i]);
      ++k1;
    }
  }
  std::vector<int> implicit;
  for (size_t k = 0; k < pairs.size(); ++k) {
    if (orders[k] != 4) continue;
    const int ends[2] = {remap[pairs[k].first], remap[pairs[k].second]};
    for (int e : ends)
      if (std::strcmp(f_type[e - 1], "C") != 0 && f_hs[e - 1] != 0 && std::find(implicit.begin(), implicit.end(), e) == implicit.end())
        implicit.push_back(e);
  }
  // ---- MOL-block text (generate_smiles.py:18-105)
  static const char* const tail = "0  0  0  0  0  0  0  0  0  0  0  0\n";
  text.clear();
  text += "\n     RDKit\n\n";
  append_rjust(text, std::to_string(f_type.size()), 3);
  append_rjust(text, std::to_string(pairs.size()), 3);
  text += "  0  0  0  0  0  0  0  0999 V2000\n";
  for (size_t i = 0; i < f_type.size(); ++i) {
    append_coord(text, static_cast<double>(f_x[i]) / 60 - 1);
    append_coord(text, static_cast<double>(f_y[i]) / 60 - 1);
    text += "    0.0000 ";
    text += f_type[i];
    text.append(4 - std::strlen(f_type[i]), ' ');
    text += tail;
  }
  for (size_t k = 0; k < pairs.size(); ++k) {
    const int o = orders[k];
    append_rjust(text, std::to_string(remap[pairs[k].first]), 3);
    append_rjust(text, std::to_string(remap[pairs[k].second]), 3);
    append_rjust(text, o <= 4 ? std::to_string(o) : std::string("1"), 3);
    append_rjust(text, o <= 4 ? std::string("0") : std::string(o == 5 ? "1" : "6"), 3);
    text += "\n";
  }
  int n_chg = 0;
  std::string line;
  for (size_t i = 0; i < f_charge.size(); ++i) {
    if (f_charge[i] == 0) continue;
    ++n_chg;
    const std::string cs = std::to_string(f_charge[i]);
    append_rjust(line, std::to_string(i + 1), 4);
    line.append(4 - cs.size(), ' ');
    line += cs;
  }
  text += "M  CHG";
  append_rjust(text, std::to_string(n_chg), 3);
  text += line;
  text += "\n";
  const int n = static_cast<int>(implicit.size());
  if (n > 0) {
    text += "M  STY  " + std::to_string(n);
    for (int k = 0; k < n; ++k) text += "   " + std::to_string(k + 1) + " DAT";
    text += "\nM  SLB  " + std::to_string(n);
    for (int k = 0; k < n; ++k) text += "   " + std::to_string(k + 1) + "   " + std::to_string(k + 1);
    text += "\n";
    for (int k = 0; k < n; ++k) {
      const std::string ks = std::to_string(k + 1);
      text += "M  SAL   " + ks + "  1  " + std::to_string(implicit[k]) + "  \n";
      text += "M  SDT   " + ks + " MRV_IMPLICIT_H    \n";
      text += "M  SDD   " + ks + "     0.0000    0.0000    DA    ALL  1       1    \n";
      text += "M  SED   " + ks + " IMPL_H1\n";
    }
  }
  text += "M  END\n$$$$";
  return true;
}

}  // namespace
}  // namespace abc

extern "C" int abc_assemble_molblocks(const AbcAtomRec* atoms, int atom_cap, const AbcBondRec* bonds, int bond_cap,
                                      const int32_t* counts, int N, const double* cos_tab, const double* sin_tab, int n_omega,
                                      int n_threads, char* text, int64_t text_stride, int32_t* text_len) {
  using namespace abc;
  ABC_REQUIRE(atoms && bonds && counts && cos_tab && sin_tab && text && text_len, "abc_assemble_molblocks: null pointer");
  ABC_REQUIRE(N > 0 && atom_cap > 0 && bond_cap > 0 && n_omega > 0 && n_omega <= 256 && text_stride > 0,
              "abc_assemble_molblocks: bad sizes");
  for (int i = 0; i < N; ++i)
    ABC_REQUIRE(counts[4 * i] >= 0 && counts[4 * i] <= atom_cap && counts[4 * i + 1] >= 0 && counts[4 * i + 1] <= bond_cap,
                "abc_assemble_molblocks: image %d: %d atoms / %d bond records exceed the capacities %d / %d", i, counts[4 * i],
                counts[4 * i + 1], atom_cap, bond_cap);
  int nt = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  if (nt < 1) nt = 1;
  if (nt > N) nt = N;
  std::atomic<int> next{0};
  std::atomic<int> overflow{-1};
  auto work = [&]() {
    std::string t;
    for (int i = next.fetch_add(1); i < N; i = next.fetch_add(1)) {
      const bool ok = assemble_one(atoms + static_cast<size_t>(i) * atom_cap, counts[4 * i], bonds + static_cast<size_t>(i) * bond_cap,
                                   counts[4 * i + 1], counts[4 * i + 2], cos_tab, sin_tab, n_omega, t);
      if (!ok) {
        text_len[i] = -1;
        continue;
      }
      text_len[i] = static_cast<int32_t>(t.size());
      if (static_cast<int64_t>(t.size()) + 1 > text_stride) {
        overflow.store(i);
        continue;
      }
      std::memcpy(text + static_cast<size_t>(i) * text_stride, t.c_str(), t.size() + 1);
    }
  };
  if (nt == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int k = 0; k < nt; ++k) pool.emplace_back(work);
    for (auto& th : pool) th.join();
  }
  if (overflow.load() >= 0) {
    set_error("abc_assemble_molblocks: image %d needs %d bytes of text, text_stride is %lld", overflow.load(),
              text_len[overflow.load()] + 1, static_cast<long long>(text_stride));
    return ABC_ERR_CAPACITY;
  }
  return ABC_OK;
}
