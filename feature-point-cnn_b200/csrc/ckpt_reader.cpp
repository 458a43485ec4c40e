// See ckpt_reader.h.  ZIP (STORED) walker + minimal pickle VM for torch.save archives.
#include "ckpt_reader.h"

#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>

namespace spb200 {
namespace {

// ------------------------------------------------------------------------------------------------
// ZIP
// ------------------------------------------------------------------------------------------------
struct ZipEntry {
    uint64_t data_offset = 0;
    uint64_t size = 0;
    bool stored = true;          // false: deflated (TorchScript code files are); such an entry is listed but never read
};

struct Archive {
    std::vector<uint8_t> bytes;
    std::map<std::string, ZipEntry> entries;
};

uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
uint64_t rd64(const uint8_t* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

void fail(const std::string& m) { throw std::runtime_error(m); }

void open_archive(const std::string& path, Archive& ar) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) fail("cannot open " + path);
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    if (sz < 22) { std::fclose(f); fail("file too small to be a torch.save archive: " + path); }
    ar.bytes.resize((size_t)sz);
    size_t got = std::fread(ar.bytes.data(), 1, (size_t)sz, f);
    std::fclose(f);
    if (got != (size_t)sz) fail("short read on " + path);
    const uint8_t* b = ar.bytes.data();
    const uint64_t n = ar.bytes.size();

    // End-of-central-directory record: scan backwards for PK\5\6.
    int64_t eocd = -1;
    for (int64_t i = (int64_t)n - 22; i >= 0 && i >= (int64_t)n - 22 - 65536; --i)
        if (rd32(b + i) == 0x06054b50u) { eocd = i; break; }
    if (eocd < 0) fail("not a ZIP archive (legacy torch.save format is not supported): " + path);
    // every offset below comes from the file: compared as `x > n - k` (never `x + k > n`, which wraps for hostile 64-bit values)
    uint64_t count = rd16(b + eocd + 10);
    uint64_t cd_off = rd32(b + eocd + 16);
    if (count == 0xFFFF || cd_off == 0xFFFFFFFFu) {   // ZIP64
        if (eocd < 20 || rd32(b + eocd - 20) != 0x07064b50u) fail("ZIP64 locator missing");
        uint64_t e64 = rd64(b + eocd - 20 + 8);
        if (n < 56 || e64 > n - 56 || rd32(b + e64) != 0x06064b50u) fail("bad ZIP64 end record");
        count = rd64(b + e64 + 32);
        cd_off = rd64(b + e64 + 48);
    }
    if (count > n / 46) fail("bad ZIP entry count");
    uint64_t p = cd_off;
    for (uint64_t k = 0; k < count; ++k) {
        if (n < 46 || p > n - 46 || rd32(b + p) != 0x02014b50u) fail("bad ZIP central directory");
        uint16_t method = rd16(b + p + 10);
        uint64_t csize = rd32(b + p + 20), usize = rd32(b + p + 24);
        uint16_t nlen = rd16(b + p + 28), xlen = rd16(b + p + 30), clen = rd16(b + p + 32);
        uint64_t lho = rd32(b + p + 42);
        if ((uint64_t)nlen + xlen + clen > n - 46 - p) fail("ZIP central directory entry runs past the end of the file");
        std::string name((const char*)b + p + 46, nlen);
        const uint8_t* x = b + p + 46 + nlen;
        const uint8_t* xe = x + xlen;
        while (xe - x >= 4) {                          // ZIP64 extended information
            uint16_t id = rd16(x), len = rd16(x + 2);
            const uint8_t* q = x + 4;
            if (len > xe - q) fail("bad ZIP extra field in " + name);
            const uint8_t* qe = q + len;
            if (id == 0x0001) {
                auto take64 = [&](uint64_t& v) { if (qe - q < 8) fail("short ZIP64 extra field in " + name); v = rd64(q); q += 8; };
                if (usize == 0xFFFFFFFFu) take64(usize);
                if (csize == 0xFFFFFFFFu) take64(csize);
                if (lho == 0xFFFFFFFFu) take64(lho);
            }
            x = qe;
        }
        if (n < 30 || lho > n - 30 || rd32(b + lho) != 0x04034b50u) fail("bad ZIP local header for " + name);
        const uint64_t hdr = 30 + (uint64_t)rd16(b + lho + 26) + rd16(b + lho + 28);
        if (hdr > n - lho) fail("bad ZIP local header for " + name);
        uint64_t data = lho + hdr;
        if (method == 0 && usize > n - data) fail("ZIP entry out of bounds: " + name);
        ar.entries[name] = ZipEntry{data, usize, method == 0};
        p += 46 + (uint64_t)nlen + xlen + clen;
    }
}

// ------------------------------------------------------------------------------------------------
// Pickle values
// ------------------------------------------------------------------------------------------------
struct PVal;
using PRef = std::shared_ptr<PVal>;

struct PVal {
    enum Kind { NONE, BOOL, INT, FLOAT, STR, TUPLE, LIST, DICT, GLOBAL, STORAGE, TENSOR, OPAQUE, MARK } kind = NONE;
    int64_t i = 0;
    double f = 0;
    std::string s, s2;                            // STR: s; GLOBAL: module s, name s2; STORAGE: dtype s, key s2
    std::vector<PRef> items;                      // TUPLE / LIST
    std::vector<std::pair<PRef, PRef>> dict;      // DICT (insertion order)
    // TENSOR
    PRef storage;
    int64_t offset = 0;
    std::vector<int64_t> sizes, strides;
};

PRef mk(PVal::Kind k) { auto v = std::make_shared<PVal>(); v->kind = k; return v; }

struct Unpickler {
    const uint8_t* p;
    const uint8_t* end;
    std::vector<PRef> stack;
    std::map<uint32_t, PRef> memo;
    uint32_t memo_next = 0;

    void need(size_t k) { if ((size_t)(end - p) < k) fail("truncated pickle"); }
    PRef pop() {
        if (stack.empty()) fail("pickle stack underflow");
        PRef v = stack.back(); stack.pop_back();
        if (!v) fail("empty value on the pickle stack");
        return v;
    }
    std::vector<PRef> pop_to_mark() {
        size_t m = stack.size();
        while (m > 0 && !(stack[m - 1] && stack[m - 1]->kind == PVal::MARK)) --m;
        if (m == 0) fail("pickle MARK not found");
        std::vector<PRef> out(stack.begin() + m, stack.end());
        stack.resize(m - 1);
        return out;
    }
    std::string line() {
        const uint8_t* q = p;
        while (q < end && *q != '\n') ++q;
        if (q == end) fail("truncated pickle line");
        std::string s((const char*)p, q - p);
        p = q + 1;
        return s;
    }
    PRef str(size_t len) {
        need(len);
        auto v = mk(PVal::STR);
        v->s.assign((const char*)p, len);
        p += len;
        return v;
    }

    static std::vector<int64_t> int_list(const PRef& t) {
        std::vector<int64_t> out;
        if (!t || (t->kind != PVal::TUPLE && t->kind != PVal::LIST)) fail("expected a tuple of ints");
        for (auto& e : t->items) {
            if (!e || e->kind != PVal::INT) fail("expected int in shape/stride");
            out.push_back(e->i);
        }
        return out;
    }

    PRef reduce(const PRef& fn, const PRef& args) {
        if (!fn || !args) fail("REDUCE on an empty value");
        for (auto& a : args->items) if (!a) fail("REDUCE argument is empty");
        if (fn->kind == PVal::GLOBAL && args->kind == PVal::TUPLE) {
            const std::string& mod = fn->s;
            const std::string& name = fn->s2;
            if (mod == "collections" && name == "OrderedDict") {
                auto d = mk(PVal::DICT);
                if (!args->items.empty() && args->items[0]->kind == PVal::LIST)   // OrderedDict([(k, v), ...])
                    for (auto& kv : args->items[0]->items)
                        if (kv && (kv->kind == PVal::TUPLE || kv->kind == PVal::LIST) && kv->items.size() == 2)
                            d->dict.emplace_back(kv->items[0], kv->items[1]);
                return d;
            }
            if (mod == "torch._utils" && (name == "_rebuild_tensor_v2" || name == "_rebuild_tensor")) {
                if (args->items.size() < 4) fail("_rebuild_tensor: too few arguments");
                auto t = mk(PVal::TENSOR);
                t->storage = args->items[0];
                if (t->storage->kind != PVal::STORAGE) fail("_rebuild_tensor: first argument is not a storage");
                if (args->items[1]->kind != PVal::INT) fail("_rebuild_tensor: bad offset");
                t->offset = args->items[1]->i;
                t->sizes = int_list(args->items[2]);
                t->strides = int_list(args->items[3]);
                if (t->sizes.size() != t->strides.size() || t->sizes.size() > 8) fail("_rebuild_tensor: sizes and strides do not match");
                int64_t numel = 1;
                for (int64_t d : t->sizes) {
                    if (d < 0 || (d > 0 && numel > (int64_t)1 << 40)) fail("_rebuild_tensor: bad size");
                    numel *= d;
                }
                if (numel > (int64_t)1 << 32 || t->offset < 0) fail("_rebuild_tensor: tensor too large or negative offset");
                return t;
            }
            if (mod == "torch._utils" && name == "_rebuild_parameter" && !args->items.empty())
                return args->items[0];
        }
        return mk(PVal::OPAQUE);
    }

    PRef persistent(const PRef& pid) {
        // ('storage', <global torch.FloatStorage>, key, location, numel)
        if (!pid || pid->kind != PVal::TUPLE || pid->items.size() < 5) fail("unsupported persistent id in pickle");
        for (auto& it : pid->items) if (!it) fail("unsupported persistent id in pickle");
        if (pid->items[0]->kind != PVal::STR || pid->items[0]->s != "storage") fail("unsupported persistent id in pickle");
        auto st = mk(PVal::STORAGE);
        const PRef& ty = pid->items[1];
        if (ty->kind == PVal::GLOBAL) st->s = ty->s2;       // e.g. FloatStorage
        else fail("storage type is not a global");
        if (pid->items[2]->kind != PVal::STR) fail("storage key is not a string");
        st->s2 = pid->items[2]->s;
        st->i = pid->items[4]->kind == PVal::INT ? pid->items[4]->i : -1;
        return st;
    }

    PRef run() {
        for (;;) {
            need(1);
            uint8_t op = *p++;
            switch (op) {
                case 0x80: need(1); ++p; break;                                       // PROTO
                case 0x95: need(8); p += 8; break;                                    // FRAME
                case '.': return pop();                                               // STOP
                case '(': stack.push_back(mk(PVal::MARK)); break;                     // MARK
                case '}': stack.push_back(mk(PVal::DICT)); break;                     // EMPTY_DICT
                case ']': stack.push_back(mk(PVal::LIST)); break;                     // EMPTY_LIST
                case ')': stack.push_back(mk(PVal::TUPLE)); break;                    // EMPTY_TUPLE
                case 'N': stack.push_back(mk(PVal::NONE)); break;                     // NONE
                case 0x88: case 0x89: { auto v = mk(PVal::BOOL); v->i = op == 0x88; stack.push_back(v); break; }
                case 'K': { need(1); auto v = mk(PVal::INT); v->i = *p; p += 1; stack.push_back(v); break; }      // BININT1
                case 'M': { need(2); auto v = mk(PVal::INT); v->i = rd16(p); p += 2; stack.push_back(v); break; } // BININT2
                case 'J': { need(4); auto v = mk(PVal::INT); v->i = (int32_t)rd32(p); p += 4; stack.push_back(v); break; }  // BININT
                case 0x8a: {                                                          // LONG1
                    need(1); uint8_t n = *p++; need(n);
                    if (n > 8) fail("LONG1 wider than 64 bits");
                    uint64_t u = 0;
                    for (int k = 0; k < n; ++k) u |= (uint64_t)p[k] << (8 * k);
                    if (n > 0 && n < 8 && (p[n - 1] & 0x80)) u |= ~0ull << (8 * n);
                    p += n;
                    auto v = mk(PVal::INT); v->i = (int64_t)u; stack.push_back(v); break;
                }
                case 'G': {                                                           // BINFLOAT (big endian)
                    need(8); uint64_t u = 0;
                    for (int k = 0; k < 8; ++k) u = (u << 8) | p[k];
                    p += 8;
                    auto v = mk(PVal::FLOAT); std::memcpy(&v->f, &u, 8); stack.push_back(v); break;
                }
                case 'X': { need(4); uint32_t n = rd32(p); p += 4; stack.push_back(str(n)); break; }   // BINUNICODE
                case 0x8c: { need(1); uint8_t n = *p++; stack.push_back(str(n)); break; }             // SHORT_BINUNICODE
                case 0x8d: { need(8); uint64_t n = rd64(p); p += 8; stack.push_back(str((size_t)n)); break; }   // BINUNICODE8
                case 'T': { need(4); uint32_t n = rd32(p); p += 4; stack.push_back(str(n)); break; }   // BINSTRING
                case 'U': { need(1); uint8_t n = *p++; stack.push_back(str(n)); break; }              // SHORT_BINSTRING
                case 'B': { need(4); uint32_t n = rd32(p); p += 4; stack.push_back(str(n)); break; }   // BINBYTES
                case 'C': { need(1); uint8_t n = *p++; stack.push_back(str(n)); break; }              // SHORT_BINBYTES
                case 'c': { auto v = mk(PVal::GLOBAL); v->s = line(); v->s2 = line(); stack.push_back(v); break; }   // GLOBAL
                case 0x93: {                                                          // STACK_GLOBAL
                    PRef name = pop(), mod = pop();
                    if (name->kind != PVal::STR || mod->kind != PVal::STR) fail("STACK_GLOBAL needs two strings");
                    auto v = mk(PVal::GLOBAL); v->s = mod->s; v->s2 = name->s; stack.push_back(v); break;
                }
                case 'q': { need(1); if (stack.empty()) fail("BINPUT on empty stack"); memo[*p] = stack.back(); p += 1; break; }           // BINPUT
                case 'r': { need(4); if (stack.empty()) fail("LONG_BINPUT on empty stack"); memo[rd32(p)] = stack.back(); p += 4; break; } // LONG_BINPUT
                case 0x94: if (stack.empty()) fail("MEMOIZE on empty stack"); memo[memo_next++] = stack.back(); break;                       // MEMOIZE
                case 'h': { need(1); auto it = memo.find(*p); p += 1; if (it == memo.end()) fail("bad memo get"); stack.push_back(it->second); break; }
                case 'j': { need(4); auto it = memo.find(rd32(p)); p += 4; if (it == memo.end()) fail("bad memo get"); stack.push_back(it->second); break; }
                case 't': { auto v = mk(PVal::TUPLE); v->items = pop_to_mark(); stack.push_back(v); break; }       // TUPLE
                case 0x85: { auto v = mk(PVal::TUPLE); v->items = {pop()}; stack.push_back(v); break; }           // TUPLE1
                case 0x86: { auto v = mk(PVal::TUPLE); PRef b = pop(), a = pop(); v->items = {a, b}; stack.push_back(v); break; }
                case 0x87: { auto v = mk(PVal::TUPLE); PRef c = pop(), b = pop(), a = pop(); v->items = {a, b, c}; stack.push_back(v); break; }
                case 'l': { auto v = mk(PVal::LIST); v->items = pop_to_mark(); stack.push_back(v); break; }        // LIST
                case 'a': { PRef x = pop(); if (stack.empty()) fail("APPEND on empty stack"); if (stack.back()->kind == PVal::LIST) stack.back()->items.push_back(x); break; }
                case 'e': { auto xs = pop_to_mark(); if (stack.empty()) fail("APPENDS on empty stack"); if (stack.back()->kind == PVal::LIST) for (auto& x : xs) stack.back()->items.push_back(x); break; }
                case 'd': {                                                           // DICT
                    auto xs = pop_to_mark(); auto v = mk(PVal::DICT);
                    for (size_t k = 0; k + 1 < xs.size(); k += 2) v->dict.emplace_back(xs[k], xs[k + 1]);
                    stack.push_back(v); break;
                }
                case 's': {                                                           // SETITEM
                    PRef val = pop(), key = pop();
                    if (stack.empty()) fail("SETITEM on empty stack");
                    if (stack.back()->kind == PVal::DICT) stack.back()->dict.emplace_back(key, val);
                    break;
                }
                case 'u': {                                                           // SETITEMS
                    auto xs = pop_to_mark();
                    if (stack.empty()) fail("SETITEMS on empty stack");
                    if (stack.back()->kind == PVal::DICT)
                        for (size_t k = 0; k + 1 < xs.size(); k += 2) stack.back()->dict.emplace_back(xs[k], xs[k + 1]);
                    break;
                }
                case 'R': { PRef args = pop(), fn = pop(); stack.push_back(reduce(fn, args)); break; }             // REDUCE
                case 'Q': { PRef pid = pop(); stack.push_back(persistent(pid)); break; }                           // BINPERSID
                case 'b': pop(); break;                                               // BUILD: state ignored (e.g. OrderedDict._metadata)
                case 0x81: { pop(); pop(); stack.push_back(mk(PVal::OPAQUE)); break; }                             // NEWOBJ
                case 0x92: { pop(); pop(); pop(); stack.push_back(mk(PVal::OPAQUE)); break; }                      // NEWOBJ_EX
                case '0': pop(); break;                                               // POP
                case '2': { if (stack.empty()) fail("DUP on empty stack"); PRef v = stack.back(); stack.push_back(v); break; }   // DUP
                case 0x8f: stack.push_back(mk(PVal::OPAQUE)); break;                  // EMPTY_SET
                case 0x90: pop_to_mark(); break;                                      // ADDITEMS
                default: {
                    char buf[64];
                    std::snprintf(buf, sizeof buf, "unsupported pickle opcode 0x%02x", op);
                    fail(buf);
                }
            }
        }
    }
};

PRef dict_get(const PRef& d, const char* key) {
    if (!d || d->kind != PVal::DICT) return nullptr;
    for (auto& kv : d->dict)
        if (kv.first && kv.first->kind == PVal::STR && kv.first->s == key) return kv.second;
    return nullptr;
}

float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ff, u;
    if (exp == 0) {
        if (man == 0) u = sign;
        else { int e = -1; do { man <<= 1; ++e; } while (!(man & 0x400)); u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13); }
    } else if (exp == 31) u = sign | 0x7f800000u | (man << 13);
    else u = sign | ((exp + 112) << 23) | (man << 13);
    float f; std::memcpy(&f, &u, 4); return f;
}

void materialise(const Archive& ar, const std::string& root, const PVal& t, HostTensor& out) {
    if (!t.storage || t.storage->kind != PVal::STORAGE || t.sizes.size() != t.strides.size()) fail("malformed tensor record");
    const PVal& st = *t.storage;
    int esize; int kind;     // kind: 0 f32, 1 f64, 2 f16, 3 bf16, 4 i64, 5 i32, 6 u8/bool
    if (st.s == "FloatStorage") { esize = 4; kind = 0; }
    else if (st.s == "DoubleStorage") { esize = 8; kind = 1; }
    else if (st.s == "HalfStorage") { esize = 2; kind = 2; }
    else if (st.s == "BFloat16Storage") { esize = 2; kind = 3; }
    else if (st.s == "LongStorage") { esize = 8; kind = 4; }
    else if (st.s == "IntStorage") { esize = 4; kind = 5; }
    else if (st.s == "ByteStorage" || st.s == "BoolStorage") { esize = 1; kind = 6; }
    else { fail("unsupported storage type " + st.s); return; }
    auto it = ar.entries.find(root + "data/" + st.s2);
    if (it == ar.entries.end()) fail("storage " + st.s2 + " missing from archive");
    if (!it->second.stored) fail("compressed storage " + st.s2 + " (torch.save writes STORED entries)");
    const uint8_t* base = ar.bytes.data() + it->second.data_offset;
    const int64_t avail = (int64_t)(it->second.size / esize);
    out.shape = t.sizes;
    out.is_integer = kind >= 4;
    const int64_t n = out.numel();
    out.data.resize((size_t)n);
    const int rank = (int)t.sizes.size();
    std::vector<int64_t> idx(rank, 0);
    for (int64_t k = 0; k < n; ++k) {
        int64_t off = t.offset;
        for (int d = 0; d < rank; ++d) off += idx[d] * t.strides[d];
        if (off < 0 || off >= avail) fail("tensor element outside its storage");
        const uint8_t* q = base + off * esize;
        float v;
        switch (kind) {
            case 0: std::memcpy(&v, q, 4); break;
            case 1: { double d; std::memcpy(&d, q, 8); v = (float)d; break; }
            case 2: v = half_to_float(rd16(q)); break;
            case 3: { uint32_t u = (uint32_t)rd16(q) << 16; std::memcpy(&v, &u, 4); break; }
            case 4: v = (float)(int64_t)rd64(q); break;
            case 5: v = (float)(int32_t)rd32(q); break;
            default: v = (float)*q; break;
        }
        out.data[(size_t)k] = v;
        for (int d = rank - 1; d >= 0; --d) { if (++idx[d] < t.sizes[d]) break; idx[d] = 0; }
    }
}

}  // namespace

bool read_checkpoint(const std::string& path, StateDict& out, std::string& err) {
    try {
        Archive ar;
        open_archive(path, ar);
        std::string root;
        bool found = false;
        for (auto& e : ar.entries) {
            const std::string& nm = e.first;
            const std::string tail = "data.pkl";
            if (nm.size() >= tail.size() && nm.compare(nm.size() - tail.size(), tail.size(), tail) == 0 &&
                (nm.size() == tail.size() || nm[nm.size() - tail.size() - 1] == '/')) {
                root = nm.substr(0, nm.size() - tail.size());
                found = true;
                break;
            }
        }
        if (!found) fail("no data.pkl in archive " + path);
        // a TorchScript archive (InferenceWrapper.trace's <name>_script.pt, python/src/inferencewrapper.py:85-87) carries
        // the model's CODE next to data.pkl; this engine executes no TorchScript, its weights come from <name>_params.pt
        for (auto& e : ar.entries)
            if (e.first.compare(0, root.size() + 5, root + "code/") == 0 || e.first == root + "constants.pkl")
                fail("'" + path + "' is a TorchScript archive: TorchScript modules are not executed by this engine; load the "
                     "weights file written next to it by InferenceWrapper.trace (<name>_params.pt), or the training checkpoint");
        auto bo = ar.entries.find(root + "byteorder");
        if (bo != ar.entries.end()) {
            std::string s;
            if (bo->second.stored) s.assign((const char*)ar.bytes.data() + bo->second.data_offset, bo->second.size);
            if (s.find("little") == std::string::npos) fail("big-endian checkpoints are not supported");
        }
        const ZipEntry& pk = ar.entries[root + "data.pkl"];
        if (!pk.stored) fail("compressed data.pkl (torch.save writes STORED entries)");
        Unpickler up{ar.bytes.data() + pk.data_offset, ar.bytes.data() + pk.data_offset + pk.size, {}, {}, 0};
        PRef top = up.run();
        if (!top || top->kind != PVal::DICT) fail("checkpoint top-level object is not a dict");
        PRef sd = dict_get(top, "model_state_dict");
        if (!sd) sd = top;
        if (sd->kind != PVal::DICT) fail("model_state_dict is not a dict");
        out.clear();
        for (auto& kv : sd->dict) {
            if (!kv.first || kv.first->kind != PVal::STR || !kv.second || kv.second->kind != PVal::TENSOR) continue;
            materialise(ar, root, *kv.second, out[kv.first->s]);
        }
        if (out.empty()) fail("no tensors found in " + path);
        return true;
    } catch (const std::exception& e) {
        err = e.what();
        return false;
    }
}

void restore_module_prefixes(StateDict& sd) {
    for (auto& kv : sd)
        if (kv.first.compare(0, 8, "encoder.") == 0 || kv.first.compare(0, 9, "detector.") == 0 || kv.first.compare(0, 11, "descriptor.") == 0)
            return;                                        // a full state_dict
    static const struct { const char* first; const char* module; } kMap[] = {
        {"conv1", "encoder."}, {"bn1", "encoder."}, {"layer1", "encoder."}, {"layer2", "encoder."},
        {"layer", "detector."},
        {"layer_in", "descriptor."}, {"up_sample", "descriptor."}, {"bn", "descriptor."}, {"layer_out", "descriptor."}};
    StateDict out;
    for (auto& kv : sd) {
        const std::string head = kv.first.substr(0, kv.first.find('.'));
        std::string key = kv.first;
        for (auto& m : kMap)
            if (head == m.first) { key = std::string(m.module) + kv.first; break; }
        out[key] = std::move(kv.second);
    }
    sd = std::move(out);
}

}  // namespace spb200
