// JSON wire format of operator trees on the C side of the boundary.
//
// The reference serialises trees with per-type to_Dict / from_Dict (src/ASTSerialization.jl:6-58 and ParDFT.jl:34-49, ParIdentity.jl:24-33,
// ParRestriction.jl:45-55, ParMatrix.jl:62-90, ParDiagonal.jl:33-60, ParTensor.jl:119-151, ParAdjoint.jl:23-28, ParCompose.jl:100-105,
// ParKron.jl:311-322).  po_plan_create_json reads exactly those dictionaries -- plus the distributed node types the reference's registry
// leaves commented out (src/ASTSerialization.jl:17,21,25):
//     ParRepartition {"type","T","global_size":[..],"dims_in":[..],"dims_out":[..],"rank":r}     ("rank" optional: the context's communicator rank)
//     ParBroadcasted {"type","of":<op>,"root":r}                                                   (apply delegates, src/ParBroadcasted.jl:33-35)
//     ParDistributed {"type","of":<op>,"rank":r,"size":p}                                          (identity only, src/ParIdentity.jl:20-22)
// and lowers them to the same flat po_node tree the hosts build, so a Julia-free harness (or a persisted plan) needs no host mirror at all.
// Parametric leaves (ParMatrix / ParDiagonal / ParTensor) get one parameter slot per distinct "id", numbered in order of first appearance
// in a depth-first walk ("of" lists left to right); po_plan_param_id names the id of a slot.
// ParAdjoint is pushed down like the reference's adjoint methods do: leaves flip, ParCompose reverses its children (src/ParCompose.jl:42),
// ParKron adjoints its factors and REVERSES its application order (src/ParKron.jl:82).  A ParKron dictionary without an "order" key gets
// the order the constructor's heuristic would pick (src/ParKron.jl:28-57).
#pragma once
#include <algorithm>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "po_comm.cuh"

// ---- a small JSON reader (objects, arrays, strings, numbers, true / false / null) -----------------------------------------------
struct po_jval {
  enum kind_t { NUL, BOOL, NUM, STR, ARR, OBJ } kind = NUL;
  double num = 0;
  bool b = false;
  std::string str;
  std::vector<po_jval> arr;
  std::vector<std::pair<std::string, po_jval> > obj;
  const po_jval* get(const char* key) const {
    for (auto& kv : obj) if (kv.first == key) return &kv.second;
    return nullptr;
  }
};

struct po_jparser {
  const char* p;
  const char* end;
  std::string err;
  int depth = 0;
  void ws() { while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
  bool fail(const char* m) { if (err.empty()) err = m; return false; }
  bool parse_string(std::string& out) {
    if (p >= end || *p != '"') return fail("expected string");
    ++p;
    out.clear();
    while (p < end && *p != '"') {
      if (*p == '\\') {
        if (++p >= end) return fail("bad escape");
        switch (*p) {
          case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break; case 'b': out += '\b'; break; case 'f': out += '\f'; break;
          case 'u': {
            if (end - p < 5) return fail("bad \\u escape");
            unsigned v = 0;
            for (int i = 1; i <= 4; ++i) {
              const char c = p[i];
              v = v * 16 + (unsigned)(c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : 0);
            }
            if (v < 0x80) out += (char)v;  // ids are ASCII in practice; keep other code points as UTF-8
            else if (v < 0x800) { out += (char)(0xC0 | (v >> 6)); out += (char)(0x80 | (v & 0x3F)); }
            else { out += (char)(0xE0 | (v >> 12)); out += (char)(0x80 | ((v >> 6) & 0x3F)); out += (char)(0x80 | (v & 0x3F)); }
            p += 4;
            break;
          }
          default: out += *p;
        }
        ++p;
      } else {
        out += *p++;
      }
    }
    if (p >= end) return fail("unterminated string");
    ++p;
    return true;
  }
  bool parse(po_jval& v) {
    if (++depth > 128) return fail("nesting too deep");
    ws();
    if (p >= end) return fail("unexpected end of input");
    bool ok = true;
    if (*p == '{') {
      v.kind = po_jval::OBJ;
      ++p; ws();
      if (p < end && *p == '}') { ++p; }
      else {
        while (ok) {
          ws();
          std::string key;
          if (!parse_string(key)) { ok = false; break; }
          ws();
          if (p >= end || *p != ':') { ok = fail("expected ':'"); break; }
          ++p;
          v.obj.emplace_back(key, po_jval());
          if (!parse(v.obj.back().second)) { ok = false; break; }
          ws();
          if (p < end && *p == ',') { ++p; continue; }
          if (p < end && *p == '}') { ++p; break; }
          ok = fail("expected ',' or '}'");
        }
      }
    } else if (*p == '[') {
      v.kind = po_jval::ARR;
      ++p; ws();
      if (p < end && *p == ']') { ++p; }
      else {
        while (ok) {
          v.arr.emplace_back();
          if (!parse(v.arr.back())) { ok = false; break; }
          ws();
          if (p < end && *p == ',') { ++p; continue; }
          if (p < end && *p == ']') { ++p; break; }
          ok = fail("expected ',' or ']'");
        }
      }
    } else if (*p == '"') {
      v.kind = po_jval::STR;
      ok = parse_string(v.str);
    } else if (end - p >= 4 && !strncmp(p, "true", 4)) { v.kind = po_jval::BOOL; v.b = true; p += 4; }
    else if (end - p >= 5 && !strncmp(p, "false", 5)) { v.kind = po_jval::BOOL; v.b = false; p += 5; }
    else if (end - p >= 4 && !strncmp(p, "null", 4)) { v.kind = po_jval::NUL; p += 4; }
    else {
      char* q = nullptr;
      v.num = strtod(p, &q);
      if (q == p || q > end) ok = fail("unexpected character");
      else { v.kind = po_jval::NUM; p = q; }
    }
    --depth;
    return ok;
  }
};

// ---- operator tree built from the dictionaries ---------------------------------------------------------------------------------
struct po_jop {
  std::string type;
  int dt = 0;                        // "T" as po_dtype (domain type of the un-adjointed operator)
  i64 n = 0, m = 0;                  // ParDFT / ParIdentity / ParRestriction / ParMatrix / ParDiagonal extents
  std::vector<i64> ranges;           // ParRestriction: s1, e1, s2, e2, ...
  std::vector<i64> ws;               // ParTensor weight shape
  i64 t_in = 0, t_out = 0;           // ParTensor prod(input_shape), prod(target_shape)
  std::string id;
  std::vector<i64> gs, dims_in, dims_out;  // ParRepartition
  i64 rank = -1, size = 1, root = 0;
  std::vector<std::unique_ptr<po_jop> > kids;
  std::vector<int> order;            // ParKron, 1-based
};

static int po_j_dtype(const po_jval& d, int& dt) {
  const po_jval* t = d.get("T");
  if (!t || t->kind != po_jval::STR) return po_fail(PO_ERR_INVALID, "dict has no 'T' key");
  if (t->str == "Float32") dt = PO_F32; else if (t->str == "Float64") dt = PO_F64; else if (t->str == "ComplexF32") dt = PO_C64; else if (t->str == "ComplexF64") dt = PO_C128;
  else return po_fail(PO_ERR_DTYPE, "unknown data type `%s`", t->str.c_str());   // src/ParDFT.jl:38-40
  return PO_OK;
}
static int po_j_int(const po_jval& d, const char* key, i64& out, bool required = true) {
  const po_jval* v = d.get(key);
  if (!v) return required ? po_fail(PO_ERR_INVALID, "dict has no '%s' key", key) : PO_OK;
  if (v->kind != po_jval::NUM || v->num != (double)(i64)v->num) return po_fail(PO_ERR_INVALID, "'%s' must be an integer", key);
  out = (i64)v->num;
  return PO_OK;
}
static int po_j_ints(const po_jval& d, const char* key, std::vector<i64>& out) {
  const po_jval* v = d.get(key);
  if (!v || v->kind != po_jval::ARR) return po_fail(PO_ERR_INVALID, "dict has no '%s' list", key);
  out.clear();
  for (auto& e : v->arr) {
    if (e.kind != po_jval::NUM || e.num != (double)(i64)e.num) return po_fail(PO_ERR_INVALID, "'%s' must be a list of integers", key);
    out.push_back((i64)e.num);
  }
  return PO_OK;
}

static int po_j_build(const po_jval& d, std::unique_ptr<po_jop>& out, int depth);

static int po_j_children(const po_jval& d, po_jop& op, int depth, bool list) {
  const po_jval* of = d.get("of");
  if (!of) return po_fail(PO_ERR_INVALID, "%s: dict has no 'of' key", op.type.c_str());
  if (list) {
    if (of->kind != po_jval::ARR || of->arr.empty()) return po_fail(PO_ERR_INVALID, "%s: 'of' must be a non-empty list", op.type.c_str());
    for (auto& c : of->arr) {
      op.kids.emplace_back();
      if (int rc = po_j_build(c, op.kids.back(), depth + 1)) return rc;
    }
    return PO_OK;
  }
  op.kids.emplace_back();
  return po_j_build(*of, op.kids.back(), depth + 1);
}

static int po_j_build(const po_jval& d, std::unique_ptr<po_jop>& out, int depth) {
  if (depth > 64) return po_fail(PO_ERR_INVALID, "operator tree too deep");
  if (d.kind != po_jval::OBJ) return po_fail(PO_ERR_INVALID, "an operator must be a JSON object");
  const po_jval* ty = d.get("type");
  if (!ty || ty->kind != po_jval::STR) return po_fail(PO_ERR_INVALID, "dict has no 'type' key");   // src/ASTSerialization.jl:38-40
  out.reset(new po_jop());
  po_jop& op = *out;
  op.type = ty->str;
  int rc = PO_OK;
  if (op.type == "ParAdjoint" || op.type == "ParBroadcasted" || op.type == "ParDistributed") {
    if ((rc = po_j_children(d, op, depth, false))) return rc;
    if (op.type == "ParBroadcasted") rc = po_j_int(d, "root", op.root, false);
    if (op.type == "ParDistributed") { if ((rc = po_j_int(d, "rank", op.rank)) || (rc = po_j_int(d, "size", op.size))) return rc; }
    return rc;
  }
  if (op.type == "ParCompose") return po_j_children(d, op, depth, true);
  if (op.type == "ParKron") {
    if ((rc = po_j_children(d, op, depth, true))) return rc;
    if (!d.get("order")) return PO_OK;  // derived by the constructor's heuristic at lowering time
    std::vector<i64> o;
    if ((rc = po_j_ints(d, "order", o))) return rc;
    if (o.size() != op.kids.size()) return po_fail(PO_ERR_INVALID, "ParKron: 'order' must list every factor");
    for (i64 v : o) op.order.push_back((int)v);
    return PO_OK;
  }
  if ((rc = po_j_dtype(d, op.dt))) return rc;
  if (op.type == "ParDFT") {
    if ((rc = po_j_int(d, "n", op.n))) return rc;
    op.m = po_dtype_is_complex(op.dt) ? op.n : op.n / 2 + 1;
    i64 m = op.m;
    if ((rc = po_j_int(d, "m", m, false))) return rc;
    if (m != op.m) return po_fail(PO_ERR_SHAPE, "range mismatch: data says %lld, but parDFT calculated %lld", (long long)m, (long long)op.m);  // src/ParDFT.jl:44-46
    return PO_OK;
  }
  if (op.type == "ParIdentity") return po_j_int(d, "n", op.n);
  if (op.type == "ParRestriction") {
    if ((rc = po_j_int(d, "n", op.n))) return rc;
    const po_jval* r = d.get("ranges");
    if (!r || r->kind != po_jval::ARR) return po_fail(PO_ERR_INVALID, "ParRestriction: dict has no 'ranges' list");
    for (auto& pr : r->arr) {
      if (pr.kind != po_jval::ARR || pr.arr.size() != 2 || pr.arr[0].kind != po_jval::NUM || pr.arr[1].kind != po_jval::NUM)
        return po_fail(PO_ERR_INVALID, "ParRestriction: ranges must be [start, stop] pairs");
      op.ranges.push_back((i64)pr.arr[0].num);
      op.ranges.push_back((i64)pr.arr[1].num);
    }
    return PO_OK;
  }
  auto read_id = [&]() -> int {
    const po_jval* idv = d.get("id");
    if (!idv || idv->kind != po_jval::STR) return po_fail(PO_ERR_INVALID, "%s: dict has no 'id' string", op.type.c_str());
    op.id = idv->str;
    return PO_OK;
  };
  if (op.type == "ParMatrix") {
    if ((rc = po_j_int(d, "m", op.m)) || (rc = po_j_int(d, "n", op.n))) return rc;
    return read_id();
  }
  if (op.type == "ParDiagonal") {
    if ((rc = po_j_int(d, "n", op.n))) return rc;
    return read_id();
  }
  if (op.type == "ParTensor") {
    std::vector<i64> is, ts;
    if ((rc = po_j_ints(d, "ws", op.ws)) || (rc = po_j_ints(d, "is", is)) || (rc = po_j_ints(d, "ts", ts))) return rc;
    op.t_in = op.t_out = 1;
    for (i64 v : is) op.t_in *= v;
    for (i64 v : ts) op.t_out *= v;
    return read_id();
  }
  if (op.type == "ParRepartition") {
    if ((rc = po_j_ints(d, "global_size", op.gs)) || (rc = po_j_ints(d, "dims_in", op.dims_in)) || (rc = po_j_ints(d, "dims_out", op.dims_out))) return rc;
    if (op.gs.empty() || op.gs.size() != op.dims_in.size() || op.gs.size() != op.dims_out.size()) return po_fail(PO_ERR_INVALID, "ParRepartition: dimensionality mismatch");
    return po_j_int(d, "rank", op.rank, false);
  }
  return po_fail(PO_ERR_INVALID, "cannot convert Dict to unknown type `%s`", op.type.c_str());   // src/ASTSerialization.jl:43-45
}

// ---- element types / sizes of a (possibly adjointed) subtree --------------------------------------------------------------------
struct po_jsig { int ddt, rdt; i64 dom, ran; };

// src/ParCommon.jl:27-30: real beats complex, lower precision beats higher
static int po_subset_type(int a, int b) {
  const bool dbl = po_dtype_is_double(a) && po_dtype_is_double(b), cplx = po_dtype_is_complex(a) && po_dtype_is_complex(b);
  return cplx ? (dbl ? PO_C128 : PO_C64) : (dbl ? PO_F64 : PO_F32);
}

// application-order heuristic of the ParKron constructor, src/ParKron.jl:28-57 (1-based operator indices)
static int po_kron_order(const std::vector<po_jsig>& sg, std::vector<int>& order) {
  const int N = (int)sg.size();
  int T = sg[0].ddt;
  for (int j = 1; j < N; ++j) T = po_subset_type(T, sg[(size_t)j].ddt);
  std::vector<char> used((size_t)N, 0);
  order.clear();
  for (int step = 0; step < N; ++step) {
    std::vector<int> cand;
    for (int j = 0; j < N; ++j) if (!used[(size_t)j] && sg[(size_t)j].ddt == T) cand.push_back(j);
    if (cand.empty()) return po_fail(PO_ERR_DTYPE, "ParKron: no factor accepts element type %s at this point of the application order", po_dtype_name(T));
    int Rt = sg[(size_t)cand[0]].rdt;
    for (int j : cand) Rt = po_subset_type(Rt, sg[(size_t)j].rdt);
    std::vector<int> c2;
    for (int j : cand) if (sg[(size_t)j].rdt == Rt) c2.push_back(j);
    i64 md = sg[(size_t)c2[0]].dom - sg[(size_t)c2[0]].ran;
    for (int j : c2) if (sg[(size_t)j].dom - sg[(size_t)j].ran > md) md = sg[(size_t)j].dom - sg[(size_t)j].ran;
    int pick = -1;
    for (int j : c2) if (sg[(size_t)j].dom - sg[(size_t)j].ran == md) pick = j;  // right-most
    used[(size_t)pick] = 1;
    order.push_back(pick + 1);
    T = sg[(size_t)pick].rdt;
  }
  return PO_OK;
}

// ---- lowering to the flat wire format ---------------------------------------------------------------------------------------------
struct po_jlower {
  std::vector<po_node> nodes;
  std::vector<int32_t> child;
  std::vector<int64_t> aux;
  std::vector<std::string> param_ids;
  std::map<std::string, int> slot_of;
  int comm_rank = 0;

  int slot(const std::string& id) {
    auto it = slot_of.find(id);
    if (it != slot_of.end()) return it->second;
    const int s = (int)param_ids.size();
    slot_of[id] = s;
    param_ids.push_back(id);
    return s;
  }
  int add(int kind, int ddt, int rdt, int adj, int slot_, i64 dom, i64 ran, const std::vector<int64_t>& a, const std::vector<int32_t>& kids) {
    po_node nd;
    memset(&nd, 0, sizeof(nd));
    nd.kind = kind; nd.ddt = ddt; nd.rdt = rdt; nd.adjoint = adj; nd.param_slot = slot_;
    nd.child_begin = (int32_t)child.size(); nd.child_count = (int32_t)kids.size();
    nd.aux_begin = (int32_t)aux.size(); nd.aux_count = (int32_t)a.size();
    nd.domain = dom; nd.range = ran;
    child.insert(child.end(), kids.begin(), kids.end());
    aux.insert(aux.end(), a.begin(), a.end());
    nodes.push_back(nd);
    return (int)nodes.size() - 1;
  }

  // application order of a ParKron as applied: the written (or heuristic) order of the un-adjointed product, reversed for the adjoint
  int kron_applied_order(const po_jop& op, bool adj, std::vector<int>& order) {
    order = op.order;
    if (order.empty()) {
      std::vector<po_jsig> sg(op.kids.size());
      for (size_t k = 0; k < op.kids.size(); ++k) if (int rc = sig(*op.kids[k], false, sg[k])) return rc;
      if (int rc = po_kron_order(sg, order)) return rc;
    }
    if (adj) std::reverse(order.begin(), order.end());   // src/ParKron.jl:82
    return PO_OK;
  }

  // signature of `op` applied as-is (adj = false) or adjointed
  int sig(const po_jop& op, bool adj, po_jsig& s) {
    po_jsig f = {0, 0, 0, 0};
    if (op.type == "ParAdjoint") return sig(*op.kids[0], !adj, s);
    if (op.type == "ParBroadcasted") return sig(*op.kids[0], adj, s);
    if (op.type == "ParDistributed") {
      int rc = sig(*op.kids[0], adj, s);
      if (rc) return rc;
      if (op.size < 1 || op.rank < 0 || op.rank >= op.size) return po_fail(PO_ERR_INVALID, "ParDistributed: bad rank %lld of %lld", (long long)op.rank, (long long)op.size);
      s.dom = po_local_size_impl(s.dom, op.rank, op.size);
      s.ran = po_local_size_impl(s.ran, op.rank, op.size);
      return PO_OK;
    }
    if (op.type == "ParCompose") {  // D = DDT(last), R = RDT(first)  (src/ParCompose.jl:22-23)
      po_jsig a, b;
      int rc = sig(*op.kids.front(), adj, a);
      if (!rc) rc = sig(*op.kids.back(), adj, b);
      if (rc) return rc;
      if (!adj) { s.ddt = b.ddt; s.dom = b.dom; s.rdt = a.rdt; s.ran = a.ran; }
      else { s.ddt = a.ddt; s.dom = a.dom; s.rdt = b.rdt; s.ran = b.ran; }   // reversed adjoint children
      return PO_OK;
    }
    if (op.type == "ParKron") {
      std::vector<po_jsig> sg(op.kids.size());
      s.dom = s.ran = 1;
      for (size_t k = 0; k < op.kids.size(); ++k) {
        if (int rc = sig(*op.kids[k], adj, sg[k])) return rc;
        s.dom *= sg[k].dom; s.ran *= sg[k].ran;
      }
      std::vector<int> order;
      if (int rc = kron_applied_order(op, adj, order)) return rc;
      if (order.empty() || order.front() < 1 || order.front() > (int)sg.size() || order.back() < 1 || order.back() > (int)sg.size())
        return po_fail(PO_ERR_INVALID, "ParKron: order is not a permutation of 1..N");
      s.ddt = sg[(size_t)order.front() - 1].ddt;
      s.rdt = sg[(size_t)order.back() - 1].rdt;
      return PO_OK;
    }
    if (op.type == "ParDFT") { f.ddt = op.dt; f.rdt = po_dtype_complex_of(op.dt); f.dom = op.n; f.ran = op.m; }
    else if (op.type == "ParIdentity") { f.ddt = f.rdt = op.dt; f.dom = f.ran = op.n; }
    else if (op.type == "ParRestriction") {
      f.ddt = f.rdt = op.dt; f.dom = op.n; f.ran = 0;
      for (size_t r = 0; r + 1 < op.ranges.size(); r += 2) if (op.ranges[r + 1] >= op.ranges[r]) f.ran += op.ranges[r + 1] - op.ranges[r] + 1;
    }
    else if (op.type == "ParMatrix") { f.ddt = f.rdt = op.dt; f.dom = op.n; f.ran = op.m; }
    else if (op.type == "ParDiagonal") { f.ddt = f.rdt = op.dt; f.dom = f.ran = op.n; }
    else if (op.type == "ParTensor") { f.ddt = f.rdt = op.dt; f.dom = op.t_in; f.ran = op.t_out; }
    else if (op.type == "ParRepartition") {
      const int N = (int)op.gs.size();
      const i64 rk = op.rank >= 0 ? op.rank : comm_rank;
      std::vector<int32_t> di(op.dims_in.begin(), op.dims_in.end()), dq(op.dims_out.begin(), op.dims_out.end());
      std::vector<int64_t> g64(op.gs.begin(), op.gs.end());
      po_repart_info info;
      if (int rc = po_repart_build(N, g64.data(), di.data(), dq.data(), (int)rk, info)) return rc;
      f.ddt = f.rdt = op.dt; f.dom = f.ran = 1;
      for (int d = 0; d < N; ++d) { f.dom *= info.local_in[(size_t)d]; f.ran *= info.local_out[(size_t)d]; }
    }
    else return po_fail(PO_ERR_INVALID, "cannot lower type `%s`", op.type.c_str());
    if (!adj) s = f;
    else { s.ddt = f.rdt; s.rdt = f.ddt; s.dom = f.ran; s.ran = f.dom; }
    return PO_OK;
  }

  int lower(const po_jop& op, bool adj, int& idx) {
    po_jsig s;
    if (int rc = sig(op, adj, s)) return rc;
    if (op.type == "ParAdjoint") return lower(*op.kids[0], !adj, idx);
    if (op.type == "ParBroadcasted") return lower(*op.kids[0], adj, idx);
    if (op.type == "ParDistributed") {
      if (op.kids[0]->type != "ParIdentity") return po_fail(PO_ERR_UNSUPPORTED, "ParDistributed apply is only defined for ParIdentity (src/ParIdentity.jl:20-22)");
      idx = add(PO_NODE_IDENTITY, s.ddt, s.rdt, 0, -1, s.dom, s.ran, {}, {});
      return PO_OK;
    }
    if (op.type == "ParCompose") {
      std::vector<int32_t> kids;
      const size_t n = op.kids.size();
      for (size_t k = 0; k < n; ++k) {
        int ci;
        if (int rc = lower(*op.kids[adj ? n - 1 - k : k], adj, ci)) return rc;
        kids.push_back(ci);
      }
      if (n == 1) { idx = kids[0]; return PO_OK; }   // src/ParCompose.jl:12-14
      idx = add(PO_NODE_COMPOSE, s.ddt, s.rdt, 0, -1, s.dom, s.ran, {}, kids);
      return PO_OK;
    }
    if (op.type == "ParKron") {
      std::vector<int32_t> kids;
      std::vector<po_jsig> sg(op.kids.size());
      for (size_t k = 0; k < op.kids.size(); ++k) {
        int ci;
        if (int rc = lower(*op.kids[k], adj, ci)) return rc;
        kids.push_back(ci);
        if (int rc = sig(*op.kids[k], adj, sg[k])) return rc;
      }
      std::vector<int> order;
      if (int rc = kron_applied_order(op, adj, order)) return rc;
      std::vector<int64_t> a(order.begin(), order.end());
      idx = add(PO_NODE_KRON, s.ddt, s.rdt, 0, -1, s.dom, s.ran, a, kids);
      return PO_OK;
    }
    if (op.type == "ParDFT") { idx = add(PO_NODE_DFT, s.ddt, s.rdt, adj ? 1 : 0, -1, s.dom, s.ran, {op.n}, {}); return PO_OK; }
    if (op.type == "ParIdentity") { idx = add(PO_NODE_IDENTITY, s.ddt, s.rdt, 0, -1, s.dom, s.ran, {}, {}); return PO_OK; }
    if (op.type == "ParRestriction") {
      std::vector<int64_t> a = {op.n, (int64_t)(op.ranges.size() / 2)};
      a.insert(a.end(), op.ranges.begin(), op.ranges.end());
      idx = add(PO_NODE_RESTRICTION, s.ddt, s.rdt, adj ? 1 : 0, -1, s.dom, s.ran, a, {});
      return PO_OK;
    }
    if (op.type == "ParMatrix") { idx = add(PO_NODE_MATRIX, s.ddt, s.rdt, adj ? 1 : 0, slot(op.id), s.dom, s.ran, {op.m, op.n}, {}); return PO_OK; }
    if (op.type == "ParDiagonal") { idx = add(PO_NODE_DIAGONAL, s.ddt, s.rdt, adj ? 1 : 0, slot(op.id), s.dom, s.ran, {op.n}, {}); return PO_OK; }
    if (op.type == "ParTensor") {
      std::vector<int64_t> a = {(int64_t)op.ws.size()};
      a.insert(a.end(), op.ws.begin(), op.ws.end());
      idx = add(PO_NODE_TENSOR, s.ddt, s.rdt, adj ? 1 : 0, slot(op.id), s.dom, s.ran, a, {});
      return PO_OK;
    }
    if (op.type == "ParRepartition") {  // the adjoint is the same exchange with the grids swapped (src/ParRepartition.jl:200-201)
      const std::vector<i64>& di = adj ? op.dims_out : op.dims_in;
      const std::vector<i64>& dq = adj ? op.dims_in : op.dims_out;
      std::vector<int64_t> a = {(int64_t)op.gs.size()};
      a.insert(a.end(), op.gs.begin(), op.gs.end());
      a.insert(a.end(), di.begin(), di.end());
      a.insert(a.end(), dq.begin(), dq.end());
      a.push_back(op.rank >= 0 ? op.rank : comm_rank);
      idx = add(PO_NODE_REPARTITION, s.ddt, s.rdt, 0, -1, s.dom, s.ran, a, {});
      return PO_OK;
    }
    return po_fail(PO_ERR_INVALID, "cannot lower type `%s`", op.type.c_str());
  }
};
