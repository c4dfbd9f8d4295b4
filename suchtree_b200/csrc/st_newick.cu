// NEWICK text -> the reference's node arrays (host code; no device work).
//
// Replaces the dendropy calls of SuchTree.__init__ (MuchTree.pyx:138-157, 171, 182,
// 200) and its two fill passes (:171-216).  dendropy is a pure-Python, un-pinned
// third-party dependency of the reference (requirements.txt:4); loading the
// 54,327-leaf ml.tree through it takes O(10 s) and a 10^6-deep caterpillar cannot be
// loaded at all (recursive iterators).  This parser is iterative, O(n), and keeps
// the rules that DEFINE node ids bit-exact:
//   * tokens: [comments] dropped, 'quoted labels' ('' = quote), ( ) , : ; and bare
//     labels (anything without whitespace or ()[],:; that does not START with a quote);
//     underscores preserved (:141)
//   * polytomies: dendropy's deterministic resolve_polytomies() (:157) -- nodes with
//     more than two children collected in post-order, then the first two children
//     are repeatedly re-attached under a new zero-length node appended LAST
//   * ids = in-order ranks of the binarised tree (:171-180)
//   * missing / zero length -> epsilon 2.22e-16, stored as fp32 (:136, :188-194);
//     root distance = -1 (:183-186); support = float(label) or -1 (:207-210)
//   * a node with exactly one child is an error (:200)
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "st_internal.cuh"

struct st_newick {
    int64_t n_nodes = 0, n_leaves = 0;
    int32_t root = -1;
    std::vector<int32_t> parent, left, right;
    std::vector<float> distance, support;
    std::vector<int32_t> leaf_ids;       // in-order (ascending id)
    std::vector<int64_t> name_offsets;   // n_leaves + 1 offsets into names
    std::string names;                   // concatenated leaf names (no separators)
};

namespace {

struct Builder {
    std::vector<int32_t> parent;
    std::vector<int32_t> first_child, last_child, next_sibling;  // child lists as linked lists
    std::vector<int32_t> n_children;
    std::vector<int64_t> label_begin;  // offset into `labels`, -1 = no label
    std::vector<int32_t> label_len;
    std::vector<double> length;
    std::vector<uint8_t> has_length;
    std::string labels;

    int32_t new_node(int32_t par) {
        const int32_t k = int32_t(parent.size());
        parent.push_back(par);
        first_child.push_back(-1);
        last_child.push_back(-1);
        next_sibling.push_back(-1);
        n_children.push_back(0);
        label_begin.push_back(-1);
        label_len.push_back(0);
        length.push_back(0.0);
        has_length.push_back(0);
        if (par >= 0) append_child(par, k);
        return k;
    }
    void append_child(int32_t par, int32_t k) {
        if (last_child[par] < 0) first_child[par] = k;
        else next_sibling[last_child[par]] = k;
        last_child[par] = k;
        next_sibling[k] = -1;
        ++n_children[par];
    }
};

inline bool is_space(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r') || c == 0x1c || c == 0x1d || c == 0x1e || c == 0x1f; }
inline bool is_punct(unsigned char c) { return c == '(' || c == ')' || c == '[' || c == ']' || c == '\'' || c == ',' || c == ':' || c == ';'; }

// Python float() on an ASCII token: decimal forms, inf / infinity / nan, optional sign
bool parse_py_float(const char *s, size_t n, double *out) {
    if (n == 0 || n > 400) return false;
    size_t i = 0;
    if (s[i] == '+' || s[i] == '-') ++i;
    const size_t body = i;
    auto ieq = [&](const char *w) {
        size_t L = strlen(w);
        if (n - body != L) return false;
        for (size_t k = 0; k < L; ++k)
            if ((s[body + k] | 0x20) != w[k]) return false;
        return true;
    };
    bool ok = false;
    if (ieq("inf") || ieq("infinity") || ieq("nan")) ok = true;
    else {
        size_t digits = 0;
        while (i < n && s[i] >= '0' && s[i] <= '9') { ++i; ++digits; }
        if (i < n && s[i] == '.') {
            ++i;
            while (i < n && s[i] >= '0' && s[i] <= '9') { ++i; ++digits; }
        }
        if (digits == 0) return false;
        if (i < n && (s[i] == 'e' || s[i] == 'E')) {
            ++i;
            if (i < n && (s[i] == '+' || s[i] == '-')) ++i;
            size_t ed = 0;
            while (i < n && s[i] >= '0' && s[i] <= '9') { ++i; ++ed; }
            if (ed == 0) return false;
        }
        ok = (i == n);
    }
    if (!ok) return false;
    char buf[408];
    memcpy(buf, s, n);
    buf[n] = 0;
    *out = strtod(buf, nullptr);  // correctly rounded, like Python's float()
    return true;
}

int fail(const char *msg) {
    st_set_error("%s", msg);
    return ST_ERR_NOT_BINARY;
}

}  // namespace

extern "C" int st_newick_parse(const char *text, int64_t len, st_newick **out) {
    if (!out) return ST_ERR_INVALID_ARG;
    *out = nullptr;
    if (!text || len < 0) {
        st_set_error("st_newick_parse: NULL text");
        return ST_ERR_INVALID_ARG;
    }
    Builder B;
    B.new_node(-1);
    int32_t cur = 0;
    bool expect_len = false, seen = false, ended = false;
    std::string tok;
    for (int64_t i = 0; i < len && !ended;) {
        const unsigned char c = (unsigned char)text[i];
        if (is_space(c)) { ++i; continue; }
        if (c == '[') {  // comment: up to the next ']' (an unterminated one is not a comment token)
            const void *e = memchr(text + i + 1, ']', size_t(len - i - 1));
            if (e) { i = (const char *)e - text + 1; continue; }
            ++i;  // the regex skips a lone '[' (it matches no alternative)
            continue;
        }
        if (c == ']') { ++i; continue; }  // matches no alternative either
        bool quoted = false;
        tok.clear();
        if (c == '\'') {
            // '(?:[^']|'')*' -- find the closing quote, '' is an escaped quote
            int64_t j = i + 1;
            bool closed = false;
            while (j < len) {
                if (text[j] == '\'') {
                    if (j + 1 < len && text[j + 1] == '\'') { tok.push_back('\''); j += 2; continue; }
                    closed = true;
                    break;
                }
                tok.push_back(text[j]);
                ++j;
            }
            if (!closed) { ++i; continue; }  // lone quote: skipped like the regex does
            quoted = true;
            i = j + 1;
        } else if (c == '(' || c == ')' || c == ',' || c == ':' || c == ';') {
            seen = true;
            ++i;
            if (c == '(') {
                cur = B.new_node(cur);
            } else if (c == ',') {
                const int32_t p = B.parent[cur];
                if (p < 0) return fail("NEWICK: ',' outside parentheses");
                cur = B.new_node(p);
                expect_len = false;
            } else if (c == ')') {
                cur = B.parent[cur];
                if (cur < 0) return fail("NEWICK: unbalanced ')'");
                expect_len = false;
            } else if (c == ':') {
                expect_len = true;
            } else {
                ended = true;  // ';' ends the first tree
            }
            continue;
        } else {
            // bare label: a quote after the first character is an ordinary character
            int64_t j = i;
            while (j < len && !is_space((unsigned char)text[j]) &&
                   (!is_punct((unsigned char)text[j]) || (text[j] == '\'' && j > i)))
                ++j;
            tok.assign(text + i, size_t(j - i));
            i = j;
        }
        seen = true;
        // a single-character quoted token that is punctuation is still a label in the
        // Python flattener only when quoted (len(tok) > 1 there); bare punctuation never gets here
        if (expect_len) {
            double v;
            // the Python flattener hands the raw token (quotes included) to float()
            if (quoted || !parse_py_float(tok.data(), tok.size(), &v)) {
                st_set_error("NEWICK: bad branch length '%.80s'", tok.c_str());
                return ST_ERR_NOT_BINARY;
            }
            B.length[cur] = v;
            B.has_length[cur] = 1;
            expect_len = false;
        } else {
            B.label_begin[cur] = int64_t(B.labels.size());
            B.label_len[cur] = int32_t(tok.size());
            B.labels += tok;
        }
    }
    if (!seen) return fail("empty NEWICK input");
    if (cur != 0) return fail("NEWICK: unbalanced '('");

    // ---- resolve polytomies: collect in post-order first, then edit
    {
        std::vector<int32_t> order, stack;
        std::vector<uint8_t> state;
        stack.push_back(0);
        state.push_back(0);
        while (!stack.empty()) {
            const int32_t v = stack.back();
            const uint8_t st = state.back();
            if (st == 0 && B.first_child[v] >= 0) {
                state.back() = 1;
                // children pushed in reverse so that the first child is visited first
                const size_t base = stack.size();
                for (int32_t ch = B.first_child[v]; ch >= 0; ch = B.next_sibling[ch]) {
                    stack.push_back(ch);
                    state.push_back(0);
                }
                for (size_t a = base, b = stack.size() - 1; a < b; ++a, --b) std::swap(stack[a], stack[b]);
            } else {
                if (B.n_children[v] > 2) order.push_back(v);
                stack.pop_back();
                state.pop_back();
            }
        }
        for (int32_t v : order) {
            while (B.n_children[v] > 2) {
                const int32_t c0 = B.first_child[v], c1 = B.next_sibling[c0];
                // detach the first two children
                B.first_child[v] = B.next_sibling[c1];
                B.n_children[v] -= 2;
                const int32_t k = B.new_node(-1);
                B.parent[k] = v;
                B.length[k] = 0.0;
                B.has_length[k] = 1;
                B.parent[c0] = k;
                B.parent[c1] = k;
                B.first_child[k] = c0;
                B.next_sibling[c0] = c1;
                B.next_sibling[c1] = -1;
                B.last_child[k] = c1;
                B.n_children[k] = 2;
                B.append_child(v, k);  // appended at the END of v's child list
            }
        }
    }
    const int64_t n = int64_t(B.parent.size());
    if (n >= (int64_t(1) << 31) - 1) return fail("NEWICK: too many nodes");
    for (int64_t v = 0; v < n; ++v) {
        if (B.n_children[v] == 1)
            return fail("node with a single child: SuchTree requires a strictly bifurcating tree");
        if (B.n_children[v] == 0 && B.label_begin[v] < 0) return fail("leaf without a name");
    }

    st_newick *R = new (std::nothrow) st_newick();
    if (!R) return ST_ERR_NOMEM;
    R->n_nodes = n;
    R->parent.assign(n, -1);
    R->left.assign(n, -1);
    R->right.assign(n, -1);
    R->distance.assign(n, 0.f);
    R->support.assign(n, -1.f);
    // ---- in-order ranks, iteratively
    std::vector<int32_t> new_id(n, -1), order;
    order.reserve(n);
    {
        std::vector<int32_t> stack;
        std::vector<uint8_t> emit;
        stack.push_back(0);
        emit.push_back(0);
        while (!stack.empty()) {
            const int32_t v = stack.back();
            const uint8_t e = emit.back();
            stack.pop_back();
            emit.pop_back();
            if (e || B.first_child[v] < 0) {
                new_id[v] = int32_t(order.size());
                order.push_back(v);
            } else {
                const int32_t c0 = B.first_child[v], c1 = B.next_sibling[c0];
                stack.push_back(c1); emit.push_back(0);
                stack.push_back(v);  emit.push_back(1);
                stack.push_back(c0); emit.push_back(0);
            }
        }
    }
    const double eps = 2.220446049250313e-16;  // np.finfo(np.float64).eps, MuchTree.pyx:136
    R->name_offsets.push_back(0);
    for (int64_t i = 0; i < n; ++i) {
        const int32_t v = order[i];
        if (B.first_child[v] >= 0) {
            const int32_t c0 = B.first_child[v], c1 = B.next_sibling[c0];
            const int32_t l = new_id[c0], r = new_id[c1];
            R->left[i] = l;
            R->right[i] = r;
            R->parent[l] = int32_t(i);
            R->parent[r] = int32_t(i);
            if (B.label_begin[v] >= 0) {
                double s;
                if (parse_py_float(B.labels.data() + B.label_begin[v], size_t(B.label_len[v]), &s))
                    R->support[i] = float(s);
            }
        } else {
            R->leaf_ids.push_back(int32_t(i));
            R->names.append(B.labels, size_t(B.label_begin[v]), size_t(B.label_len[v]));
            R->name_offsets.push_back(int64_t(R->names.size()));
        }
        const double ln = (B.has_length[v] && B.length[v] != 0.0) ? B.length[v] : eps;  // None, 0.0, -0.0 -> eps
        R->distance[i] = float(ln);
    }
    R->n_leaves = int64_t(R->leaf_ids.size());
    R->root = new_id[0];
    R->distance[R->root] = -1.0f;
    *out = R;
    return ST_OK;
}

extern "C" void st_newick_free(st_newick *p) { delete p; }

extern "C" int st_newick_info(const st_newick *p, int64_t *n_nodes, int64_t *n_leaves, int32_t *root,
                              int64_t *names_bytes) {
    if (!p) return ST_ERR_INVALID_ARG;
    if (n_nodes) *n_nodes = p->n_nodes;
    if (n_leaves) *n_leaves = p->n_leaves;
    if (root) *root = p->root;
    if (names_bytes) *names_bytes = int64_t(p->names.size());
    return ST_OK;
}

extern "C" int st_newick_arrays(const st_newick *p, int32_t *parent, int32_t *left, int32_t *right,
                                float *distance, float *support) {
    if (!p) return ST_ERR_INVALID_ARG;
    const size_t n = size_t(p->n_nodes);
    if (parent) memcpy(parent, p->parent.data(), n * 4);
    if (left) memcpy(left, p->left.data(), n * 4);
    if (right) memcpy(right, p->right.data(), n * 4);
    if (distance) memcpy(distance, p->distance.data(), n * 4);
    if (support) memcpy(support, p->support.data(), n * 4);
    return ST_OK;
}

extern "C" int st_newick_leaves(const st_newick *p, int32_t *leaf_ids, int64_t *name_offsets, char *names) {
    if (!p) return ST_ERR_INVALID_ARG;
    if (leaf_ids) memcpy(leaf_ids, p->leaf_ids.data(), p->leaf_ids.size() * 4);
    if (name_offsets) memcpy(name_offsets, p->name_offsets.data(), p->name_offsets.size() * 8);
    if (names) memcpy(names, p->names.data(), p->names.size());
    return ST_OK;
}
