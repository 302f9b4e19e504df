// GMNS file contract + the drop-in entry point (include/dtalite_b200.h, level A).
//
//   dtl_gmns_read            <- g_read_input_data (reference input.h:2388-3795) +
//                               g_ReadDemandFileBasedOnDemandFileList, column format (input.h:392, :993-1203) +
//                               virtual connectors (main_api.cpp:534-570, input.h:1859-1908)
//   dtl_gmns_write_outputs   <- g_output_assignment_result (output.h:486-1459): link_performance.csv and
//                               route_assignment.csv with the reference's headers, column order and formats
//   network_assignment       <- main_api.cpp:499 (same seven ints; returns 1.0)
//   dtalite_main             <- flash_dta.cpp:140-246
//
// Host-only C++; the compute is behind the dtl_* engine calls.  Inputs outside the accelerated path
// (matrix/path/activity demand, subarea.csv, multimodal access links, zone generation, ODME, scenario
// analysis, simulation) stop with an explicit message and a non-zero exit status; nothing blocks on stdin
// (the reference's g_program_stop() does, utils.cpp:41-46).
#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "../../include/dtalite_b200.h"

namespace {

thread_local std::string g_io_err;

struct StopError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
[[noreturn]] void stop(const std::string& msg) { throw StopError(msg); }

// ---------------------------------------------------------------------------------------------
// CSV dialect of the reference's CDTACSVParser (utils.cpp:233-470, utils.h:115-221): header names are
// left-trimmed of blanks, later duplicates win, a line containing '[' opens a section and is its header,
// an empty line ends ReadRecord loops, quoted fields may contain commas.
// The whole file is read into memory once; a record's fields are views into that buffer (commas and the line end are
// overwritten with NULs so that strtod / strtol run on the field in place): no per-field std::string, which is what the
// reader of a 300 000-row link.csv used to spend its time on.  Lines with quotes take the general splitter.
class Csv {
public:
    bool first_line_header = true;
    std::string section;
    std::string file;

    bool open(const std::string& path)
    {
        file = path;
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) return false;
        buf_.clear();
        char chunk[1 << 16];
        size_t n;
        while ((n = fread(chunk, 1, sizeof(chunk), f)) > 0) buf_.append(chunk, n);
        fclose(f);
        buf_.push_back('\n'); // sentinel: the last line always ends inside the buffer
        pos_ = 0;
        end_ = buf_.size() - 1;
        open_ = true;
        if (first_line_header) {
            std::string s;
            getline_buf(s);
            set_header(s);
        }
        return true;
    }
    void close() { open_ = false; buf_.clear(); }

    // a slice of a file that was read elsewhere (whole lines), with that file's header line: what a worker thread parses
    void open_text(const char* begin, const char* end, const std::string& header_line, const std::string& path)
    {
        file = path;
        buf_.assign(begin, end);
        buf_.push_back('\n');
        pos_ = 0;
        end_ = buf_.size() - 1;
        open_ = true;
        set_header(header_line);
    }
    static bool read_file(const std::string& path, std::string& out)
    {
        FILE* f = fopen(path.c_str(), "rb");
        if (!f) return false;
        out.clear();
        char chunk[1 << 16];
        size_t n;
        while ((n = fread(chunk, 1, sizeof(chunk), f)) > 0) out.append(chunk, n);
        fclose(f);
        return true;
    }

    bool read_record()
    {
        fields_.clear();
        if (!open_ || pos_ >= end_) return false; // std::getline at end of file leaves the string empty
        char* b = &buf_[pos_];
        char* e = (char*)memchr(b, '\n', end_ + 1 - pos_);
        pos_ = (size_t)(e - &buf_[0]) + 1;
        if (e == b) return false;                 // an empty line ends ReadRecord loops (utils.cpp:290-312)
        if (memchr(b, '"', (size_t)(e - b))) {    // quoted fields: general splitter, fields kept in side strings
            quoted_ = split(std::string(b, e));
            for (std::string& q : quoted_) fields_.push_back({ &q[0], (int)q.size() });
            for (size_t i = 0; i < quoted_.size(); ++i)
                if (quoted_[i].empty()) fields_[i] = { empty_, 0 };
            return true;
        }
        *e = '\0';
        char* p = b;
        for (;;) {
            char* c = (char*)memchr(p, ',', (size_t)(e - p));
            if (!c) {
                if (p < e) fields_.push_back({ p, (int)(e - p) });
                break;
            }
            *c = '\0';
            fields_.push_back({ p, (int)(c - p) });
            p = c + 1;
        }
        if (e[-1] == '\0' && e > b) fields_.push_back({ empty_, 0 }); // the line ended with a comma: one more, empty, field
        return true;
    }

    // sectioned settings.csv: returns false only at end of file
    bool read_record_section()
    {
        fields_.clear();
        if (!open_) return false;
        std::string s;
        const bool more = getline_buf(s);
        if (!s.empty()) {
            if (s.find('[') != std::string::npos) {
                std::vector<std::string> v = split(s);
                if (!v.empty()) section = v[0];
                set_header(s);
                getline_buf(s);
            }
            quoted_ = split(s);
            for (std::string& q : quoted_) {
                if (q.empty()) fields_.push_back({ empty_, 0 });
                else fields_.push_back({ &q[0], (int)q.size() });
            }
            return true;
        }
        return more;
    }

    bool has(const std::string& name) const { return index_.count(name) != 0; }

    bool get(const std::string& name, std::string& value, bool required = true)
    {
        const Field* s = raw(name, required);
        if (!s) return false;
        value.assign(s->p, (size_t)s->n);
        return true;
    }
    template <class T>
    bool get(const std::string& name, T& value, bool required = true, bool nonnegative = true)
    {
        const Field* s = raw(name, required);
        if (!s) return false;
        T v;
        if (!parse(s->p, v)) return false;
        if (required && nonnegative && v < 0) v = 0;
        value = v;
        return true;
    }
    // Field names are string literals at the call sites: the literal's address caches the column index, so a row of
    // thirty fields costs thirty pointer look-ups instead of thirty std::string constructions and map searches.
    template <class T>
    bool get(const char* name, T& value, bool required = true, bool nonnegative = true)
    {
        const Field* s = raw_cached(name, required);
        if (!s) return false;
        T v;
        if (!parse(s->p, v)) return false;
        if (required && nonnegative && v < 0) v = 0;
        value = v;
        return true;
    }
    bool get(const char* name, std::string& value, bool required = true)
    {
        const Field* s = raw_cached(name, required);
        if (!s) return false;
        value.assign(s->p, (size_t)s->n);
        return true;
    }
    bool has(const char* name) const { return index_.count(name) != 0; }

private:
    struct Field { char* p; int n; }; // NUL-terminated view into buf_ (or into quoted_)
    std::string buf_;
    size_t pos_ = 0, end_ = 0;
    bool open_ = false;
    char empty_[1] = { 0 };
    std::map<std::string, int> index_;
    std::vector<Field> fields_;
    std::vector<std::string> quoted_;
    std::vector<std::pair<const char*, int>> cache_; // literal address -> column index (-1: no such column); reset with the header

    // std::getline over the buffer: false once the end of the file has been reached (the string is empty then)
    bool getline_buf(std::string& s)
    {
        s.clear();
        if (pos_ >= end_) return false;
        const char* b = &buf_[pos_];
        const char* e = (const char*)memchr(b, '\n', end_ + 1 - pos_);
        s.assign(b, e);
        pos_ = (size_t)(e - &buf_[0]) + 1;
        return pos_ < end_ || true;
    }

    // the stream extraction the reference's parser performs (operator>> of an istringstream), without the stream:
    // leading white space skipped, the longest valid prefix converted, failure when nothing converts
    static bool parse(const char* s, double& v) { char* e; v = strtod(s, &e); return e != s; }
    static bool parse(const char* s, float& v) { char* e; v = strtof(s, &e); return e != s; }
    static bool parse(const char* s, long& v)
    {
        char* e;
        errno = 0;
        v = strtol(s, &e, 10);
        return e != s && errno != ERANGE;
    }
    static bool parse(const char* s, int& v)
    {
        char* e;
        const long long x = strtoll(s, &e, 10);
        if (e == s) return false;
        if (x > 2147483647LL || x < -2147483648LL) return false; // operator>> sets failbit on overflow
        v = (int)x;
        return true;
    }
    const Field* raw_cached(const char* name, bool required)
    {
        int col = -2;
        for (const auto& c : cache_)
            if (c.first == name) { col = c.second; break; }
        if (col == -2) {
            auto it = index_.find(name);
            col = it == index_.end() ? -1 : it->second;
            cache_.emplace_back(name, col);
        }
        if (col < 0) {
            if (required) stop(std::string("Field ") + name + " in file " + file + " does not exist. Please check the file.");
            return nullptr;
        }
        if (fields_.empty() || col >= (int)fields_.size()) return nullptr;
        const Field& s = fields_[col];
        return s.n == 0 ? nullptr : &s;
    }

    void set_header(const std::string& line)
    {
        index_.clear();
        cache_.clear();
        if (line.empty()) return;
        std::vector<std::string> names = split(line);
        for (size_t i = 0; i < names.size(); ++i) {
            const size_t b = names[i].find_first_not_of(' ');
            index_[b == std::string::npos ? std::string() : names[i].substr(b)] = (int)i;
        }
    }
    const Field* raw(const std::string& name, bool required)
    {
        auto it = index_.find(name);
        if (it == index_.end()) {
            if (required) stop("Field " + name + " in file " + file + " does not exist. Please check the file.");
            return nullptr;
        }
        if (fields_.empty() || it->second >= (int)fields_.size()) return nullptr;
        const Field& s = fields_[it->second];
        return s.n == 0 ? nullptr : &s;
    }
    static std::vector<std::string> split(const std::string& line)
    {
        std::vector<std::string> out;
        if (line.empty()) return out;
        if (line.find('"') == std::string::npos) {
            size_t pos = 0;
            while (pos <= line.size()) {
                const size_t c = line.find(',', pos);
                if (c == std::string::npos) {
                    if (pos < line.size()) out.push_back(line.substr(pos));
                    break;
                }
                out.push_back(line.substr(pos, c - pos));
                pos = c + 1;
            }
            if (line.back() == ',') out.push_back("");
            return out;
        }
        size_t pos = 0; // start of the unread remainder
        while (pos < line.size()) {
            const size_t c = line.find(',', pos);
            const size_t q = line.find('"', pos);
            if (c == std::string::npos && q == std::string::npos) { out.push_back(line.substr(pos)); break; }
            if (q != std::string::npos && (c == std::string::npos || q < c)) { // quoted field
                const size_t q2 = line.find('"', q + 1);
                out.push_back(q2 == std::string::npos ? line.substr(q + 1) : line.substr(q + 1, q2 - q - 1));
                if (q2 == std::string::npos) break;
                const size_t next = line.find(',', q2 + 1);
                if (next == std::string::npos) break;
                pos = next + 1;
                continue;
            }
            out.push_back(line.substr(pos, c - pos)); // plain field
            if (c + 1 < line.size()) pos = c + 1;
            else { out.push_back(""); break; }
        }
        return out;
    }
};

// HHMM_HHMM / DDHHMM_DDHHMM -> minutes, float arithmetic as in g_time_parser (shared_code.h:76-196)
std::vector<float> parse_time_period(const std::string& s)
{
    std::vector<float> out;
    std::string tok;
    auto flush = [&]() {
        float dd = 0, hh = 0, mm = 0;
        if (tok.size() == 4) {
            hh = ((float)tok[0] - 48) * 10 * 60 + ((float)tok[1] - 48) * 60;
            mm = ((float)tok[2] - 48) * 10 + ((float)tok[3] - 48);
        } else if (tok.size() == 6) {
            dd = ((float)tok[0] - 48) * 10 * 24 * 60 + ((float)tok[1] - 48) * 24 * 60;
            hh = ((float)tok[2] - 48) * 10 * 60 + ((float)tok[3] - 48) * 60;
            mm = ((float)tok[4] - 48) * 10 + ((float)tok[5] - 48);
        }
        out.push_back(dd + hh + mm);
        tok.clear();
    };
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] == ':') stop("time_period '" + s + "': second-resolution periods are not supported");
        if (s[i] != '_') tok.push_back(s[i]);
        if (s[i] == '_' || i + 1 == s.size()) flush();
    }
    return out;
}

struct AgentType { std::string name; float vot = 100; double pce = 1, occ = 1; int rt_info = 0; std::string access_node_type; };
struct LinkType { int id = 1; std::string type_code; int vdf_type = 0 /* q_vdf, CLinkType() DTA.h:505 */; double k_jam = 300; };
struct Period { int id = 0; std::string name, time_period; int start_slot = 0, end_slot = 0; float hours = 1, t2 = 0; int n_files = 0; };

} // namespace

// ================================================================================================
struct dtl_gmns {
    dtl_problem prob{};
    std::string dir;
    std::vector<AgentType> agents;
    std::vector<Period> periods;
    // node / zone identity
    std::vector<int> node_id, node_zone_org, node_zone, zone_id, zone_centroid;
    std::vector<double> node_x, node_y;
    // link identity / attributes for the writers
    std::vector<std::string> link_id, geometry, link_type_code, vdf_code, tmc_code, tmc_corridor_name, link_allowed_uses;
    std::vector<int> link_type, link_vdf_type, meso_link_id, tmc_corridor_id, tmc_road_sequence;
    std::vector<double> free_speed, cutoff_speed, dist_km, dist_mile, fftt_min, ref_volume;
    // problem arrays
    std::vector<int> link_from, link_to, link_tag;
    std::vector<double> fftt, cap, alpha, beta, vf, vc, nlanes, cap_lane, qdf, preload, penalty, Q_alpha, Q_beta, Q_cd, Q_n, Q_cp,
        Q_s, t2, L_h, start_h, end_h, plf, pce, occ, toll, lr_price, od_vol;
    std::vector<float> cycle_length, red_time, sat_flow, vot, origin_demand;
    std::vector<int> vdf_type, od_o, od_d, od_at, od_tau;
    std::vector<unsigned char> allowed;
    int path_output = 1;
    int sa_iterations = 0;
    // supply_side_scenario.csv rows of type "sa" (scenario_API.h:152-160): lanes of (period, link) during the sensitivity analysis
    std::vector<int> sa_tau, sa_link;
    std::vector<int> dms_tau, dms_link, dms_zone; // "dms" rows: period, link, zone seq no of the link's to-node (-1: its zone id is no zone of the model)
    std::vector<double> sa_lanes;
    // measurement.csv rows of type "link" (ODME.h:100-178), one entry per (row, period) in file order
    std::vector<int> ms_tau, ms_link, ms_ub;
    std::vector<float> ms_count;
    std::vector<double> obs_count;         // [T*L] the record a link ends up with (0 = none) and its upper_bound_flag: what the writers print
    std::vector<int> obs_ub;
    std::vector<signed char> design_flag;  // [T*L] network_design_flag: -1 on the links of the sa rows
    // link state recorded before the sensitivity-analysis stage (g_record_corridor_performance_summary(0), output.h:52-98):
    // [T*L] vehicle volume, speed, DOC, P -- the *_before columns of link_performance.csv; empty = not recorded
    std::vector<double> before_volume, before_speed, before_doc, before_P;
    bool sa_stage = false;                 // the stage ran: per-column flags and per-iteration records are printed
    bool odme_ran = false;                 // ODME ran: link_performance.csv prints the count columns of the links with a record
};

namespace {

void read_settings(dtl_gmns& g)
{
    const std::string settings = g.dir + "/settings.csv";
    {   // [demand_period], input.h:2412-2508
        Csv p; p.first_line_header = false;
        if (!p.open(settings)) stop("File settings.csv cannot be opened.");
        while (p.read_record_section()) {
            if (p.section != "[demand_period]") continue;
            Period d;
            if (!p.get("demand_period_id", d.id)) break;
            if (!p.get("demand_period", d.name)) stop("Field demand_period in section [demand_period] cannot be read.");
            if (!p.get("time_period", d.time_period)) stop("Field time_period in section [demand_period] cannot be read.");
            std::vector<float> m = parse_time_period(d.time_period);
            if (m.size() == 2) {
                d.start_slot = m[0] / 5; // MIN_PER_TIMESLOT, DTA.h:29
                d.end_slot = m[1] / 5;
                d.hours = (m[1] - m[0]) / 60.0;
                d.t2 = (m[0] + m[1]) / 2 / 60;
                std::string peak;
                if (p.get("peak_time", peak, false)) stop("[demand_period] peak_time is not supported");
            }
            g.periods.push_back(d);
        }
        if (g.periods.empty()) stop("Section demand_period has no information.");
        // the reference stops at >= MAX_TIMEPERIODS (= 2, DTA.h:15) periods; this engine has no such cap
    }
    {   // [agent_type], input.h:2599-2696
        Csv p; p.first_line_header = false;
        p.open(settings);
        while (p.read_record_section()) {
            if (p.section != "[agent_type]") continue;
            AgentType a;
            if (!p.get("agent_type", a.name)) break;
            for (const AgentType& b : g.agents)
                if (b.name.find(a.name) != std::string::npos)
                    stop("agent_type " + b.name + " in section agent_type is overlapping with " + a.name);
            p.get("vot", a.vot, false, false);
            p.get("pce", a.pce, false, false);
            p.get("person_occupancy", a.occ, false, false);
            p.get("real_time_info", a.rt_info, false);
            if (a.name == "rt") a.rt_info = 1;
            if (a.name == "dms") a.rt_info = 2;
            p.get("access_node_type", a.access_node_type, false);
            if (!a.access_node_type.empty()) stop("multimodal access links (agent_type.access_node_type) are not supported");
            g.agents.push_back(a);
        }
        if (g.agents.empty()) stop("Section agent_type does not contain information.");
        // the reference stops at >= MAX_AGNETTYPES (= 5, DTA.h:14) agent types; this engine has no such cap
    }
}

std::map<int, LinkType> read_link_types(const dtl_gmns& g)
{   // [link_type], input.h:2513-2595
    std::map<int, LinkType> m;
    LinkType vc; vc.id = -1; vc.type_code = "c";
    m[-1] = vc;
    Csv p; p.first_line_header = false;
    if (!p.open(g.dir + "/settings.csv")) return m;
    int line_no = 0;
    while (p.read_record_section()) {
        if (p.section != "[link_type]") continue;
        LinkType e;
        if (!p.get("link_type", e.id)) {
            if (line_no == 0) stop("Field link_type cannot be found in file link_type section.");
            break;
        }
        if (m.count(e.id)) stop("link_type " + std::to_string(e.id) + " has been defined more than once");
        p.get("type_code", e.type_code, true);
        std::string v;
        e.vdf_type = 1; // bpr unless the section says qvdf, input.h:2559-2566
        p.get("vdf_type", v, false);
        if (v == "bpr") e.vdf_type = 1;
        if (v == "qvdf") e.vdf_type = 0;
        p.get("k_jam", e.k_jam, false);
        m[e.id] = e;
        ++line_no;
    }
    return m;
}

void read_network(dtl_gmns& g)
{
    const int AT = (int)g.agents.size(), T = (int)g.periods.size();
    std::map<int, LinkType> link_types = read_link_types(g);
    {   // zone.csv would change zone membership (access_node_vector); its production/attraction are unused here
        Csv z;
        if (z.open(g.dir + "/zone.csv")) {
            while (z.read_record()) {
                std::string s;
                if (z.has("access_node_vector") && z.get("access_node_vector", s, false))
                    stop("zone.csv access_node_vector is not supported");
            }
        }
    }
    // ---- node.csv, input.h:2803-2994 ----
    std::map<int, int> node_seq, zone_first_node;
    {
        Csv p;
        if (!p.open(g.dir + "/node.csv")) stop("File node.csv does not exist. Please check.");
        while (p.read_record()) {
            int node_id;
            if (!p.get("node_id", node_id)) continue;
            if (node_seq.count(node_id)) continue;
            double x = 0, y = 0;
            p.get("x_coord", x, true, false);
            p.get("y_coord", y, true, false);
            int zone = -1;
            p.get("zone_id", zone);
            node_seq[node_id] = (int)g.node_id.size();
            g.node_id.push_back(node_id);
            g.node_x.push_back(x); g.node_y.push_back(y);
            g.node_zone_org.push_back(zone >= 1 ? zone : -1);
            if (zone >= 1 && !zone_first_node.count(zone)) zone_first_node[zone] = node_id;
        }
    }
    if (zone_first_node.size() <= 1) stop("node.csv defines fewer than two zones (zone generation is not supported)");
    const int n_phys = (int)g.node_id.size();
    // ---- centroids in ascending zone_id, input.h:3138-3180 ----
    std::map<int, int> zone_seq;
    for (auto& kv : zone_first_node) {
        const int zid = kv.first;
        zone_seq[zid] = (int)g.zone_id.size();
        g.zone_id.push_back(zid);
        const int first = node_seq[kv.second];
        node_seq[-zid] = (int)g.node_id.size();
        g.zone_centroid.push_back((int)g.node_id.size());
        g.node_id.push_back(-zid);
        g.node_x.push_back(g.node_x[first]); g.node_y.push_back(g.node_y[first]);
        g.node_zone_org.push_back(-1);
    }
    const int N = (int)g.node_id.size();
    g.node_zone.assign(N, -1);
    for (size_t z = 0; z < g.zone_centroid.size(); ++z) g.node_zone[g.zone_centroid[z]] = (int)z;
    if (!std::ifstream(g.dir + "/" + "link.csv").good()) stop("File link.csv does not exist. Please check.");

    // per-link, per-period temporaries in link-major order; transposed to [T][L] at the end
    struct PerPeriod {
        double fftt, cap, alpha, beta, vf, vc, nlanes, cap_lane, qdf, preload, penalty, Q_alpha, Q_beta, Q_cd, Q_n, Q_cp, Q_s, t2, L,
            start_h, end_h, plf;
        float cycle, red, sat;
        int vdf_type;
        std::string allowed_uses;
        std::vector<double> at4; // [4][AT]: pce, occ, toll, lr of every agent type in one block
        int AT = 0;
        double* pce() { return at4.data(); }
        double* occ() { return at4.data() + AT; }
        double* toll() { return at4.data() + 2 * AT; }
        double* lr() { return at4.data() + 3 * AT; }
        const double* pce() const { return at4.data(); }
        const double* occ() const { return at4.data() + AT; }
        const double* toll() const { return at4.data() + 2 * AT; }
        const double* lr() const { return at4.data() + 3 * AT; }
    };
    std::vector<std::vector<PerPeriod>> lp;
    auto make_default_period = [&]() { // CPeriod_VDF(), VDF.h:57-74
        PerPeriod v{};
        v.fftt = 1; v.cap = 1; v.alpha = 0.39999993; v.beta = 4; v.vf = 60; v.vc = 45; v.nlanes = 1; v.cap_lane = 0; v.qdf = -1;
        v.preload = 0; v.penalty = 0; v.Q_alpha = 0.272876961; v.Q_beta = 4; v.Q_cd = 0.954946463; v.Q_n = 1.141574427;
        v.Q_cp = 0.400089684; v.Q_s = 4; v.t2 = 1; v.L = 1; v.start_h = 0; v.end_h = 0; v.plf = 1;
        v.cycle = -1; v.red = 0; v.sat = 1800; v.vdf_type = 0;
        v.AT = AT;
        v.at4.assign((size_t)4 * AT, 0.0);
        for (int at = 0; at < AT; ++at) { v.pce()[at] = 1.0; v.occ()[at] = 1.0; }
        return v;
    };
    const PerPeriod period_proto = make_default_period();
    auto default_period = [&](int) -> const PerPeriod& { return period_proto; };
    // ---- link.csv, input.h:3207-3678 ----
    {
        Csv p;
        // per-period / per-agent-type field names, built once: their c_str() addresses key the reader's column cache
        std::vector<std::string> n_pce(AT), n_fftt(T), n_cap(T), n_alpha(T), n_beta(T), n_uses(T), n_preload(T), n_penalty(T),
            n_toll((size_t)T * AT);
        for (int at = 0; at < AT; ++at) n_pce[at] = "VDF_pce" + g.agents[at].name;
        for (int tau = 0; tau < T; ++tau) {
            const std::string pid = std::to_string(g.periods[tau].id);
            n_fftt[tau] = "VDF_fftt" + pid; n_cap[tau] = "VDF_cap" + pid; n_alpha[tau] = "VDF_alpha" + pid; n_beta[tau] = "VDF_beta" + pid;
            n_uses[tau] = "VDF_allowed_uses" + pid; n_preload[tau] = "VDF_preload" + pid; n_penalty[tau] = "VDF_penalty" + pid;
            for (int at = 0; at < AT; ++at) n_toll[(size_t)tau * AT + at] = "VDF_toll" + g.agents[at].name + pid;
        }
        const std::string n_ref = "VDF_ref_link_volume" + std::to_string(g.periods[0].id);
        // The rows of link.csv are independent of each other except for one thing (the first ten link types that [link_type]
        // does not define are registered in row order), so the file is cut into slices of whole lines that worker threads parse
        // into their own vectors, appended in slice order afterwards; a file with an unknown link type is re-read serially.
        auto parse_rows = [&](Csv& p, dtl_gmns& out, std::vector<std::vector<PerPeriod>>& lp_out, std::map<int, LinkType>& link_types,
                              bool serial, bool& unknown_type) {
            int type_warnings = 0;
        while (p.read_record()) {
            long from_id = -1, to_id = -1;
            if (!p.get("from_node_id", from_id)) continue;
            if (!p.get("to_node_id", to_id)) continue;
            std::string id;
            p.get("link_id", id, false);
            auto fi = node_seq.find((int)from_id), ti = node_seq.find((int)to_id);
            if (fi == node_seq.end() || ti == node_seq.end()) continue; // endpoint not in node.csv: skipped with a warning there
            const int from = fi->second, to = ti->second;
            int cell_type = -1;
            p.get("cell_type", cell_type, false);
            int meso = -1;
            p.get("meso_link_id", meso, false);
            std::string geometry, type_code, vdf_code, tmc, tmc_name = "network_wide";
            p.get("geometry", geometry, false);
            p.get("tmc_corridor_name", tmc_name, false);
            p.get("link_type_code", type_code, false);
            int link_type_in = 2, link_type = 2; // CLink(): link_type 2, DTA.h:1105
            p.get("link_type", link_type_in, false);
            if (!link_types.count(link_type_in)) {
                if (!serial) { unknown_type = true; return; } // registering a type depends on the row order: the caller re-reads serially
                if (type_warnings < 10) { // the reference registers the unknown type (first ten only) but leaves the link at type 2
                    LinkType e; e.id = link_type_in;
                    link_types[link_type_in] = e;
                    ++type_warnings;
                }
            } else link_type = link_type_in;
            const LinkType& lt = link_types[link_type];
            int tag = -1;
            if (lt.type_code == "c" && g.node_zone_org[from] >= 1) {
                auto zi = zone_seq.find(g.node_zone_org[from]);
                if (zi != zone_seq.end()) tag = zi->second;
            }
            double length_m = 1.0, free_speed = 60.0, lanes = 1, lane_cap = 1800;
            float sat_in = 1800;
            p.get("length", length_m);
            p.get("vdf_code", vdf_code, false);
            if (length_m < 1) length_m = 1;
            p.get("free_speed", free_speed);
            double cutoff = free_speed * 0.85;
            p.get("cutoff_speed", cutoff, false);
            if (free_speed <= 0.1) free_speed = 60;
            free_speed = std::max(0.1, free_speed);
            p.get("lanes", lanes);
            p.get("capacity", lane_cap);
            p.get("saturation_flow_rate", sat_in, false);
            const double dist = length_m / 1609.0;
            std::vector<PerPeriod> per(T);
            std::vector<double> pce_at(AT, -1.0);
            for (int at = 0; at < AT; ++at) p.get(n_pce[at].c_str(), pce_at[at], false, true);
            for (int tau = 0; tau < T; ++tau) {
                PerPeriod v = default_period(tau);
                const Period& d = g.periods[tau];
                for (int at = 0; at < AT; ++at) {
                    v.pce()[at] = pce_at[at] > 0 ? pce_at[at] : g.agents[at].pce;
                    v.occ()[at] = g.agents[at].occ;
                }
                v.vdf_type = lt.vdf_type;
                v.cap_lane = lane_cap; v.nlanes = lanes;
                v.fftt = dist / std::max(0.0001, free_speed) * 60.0;
                v.cap = lane_cap * lanes * d.hours;
                v.vf = free_speed; v.vc = (float)cutoff /* CLink::v_congestion_cutoff is a float, DTA.h:1304 */; v.alpha = 0.15; v.beta = 4; v.preload = 0;
                v.start_h = d.start_slot * 5 / 60.0;
                v.end_h = d.end_slot * 5 / 60.0;
                v.L = d.hours; v.t2 = d.t2; v.plf = 1;
                double fftt_in = -1;
                p.get(n_fftt[tau].c_str(), fftt_in, false, false);
                const bool req = fftt_in >= 0;
                if (req) v.fftt = fftt_in;
                p.get(n_cap[tau].c_str(), v.cap, req, false);
                p.get(n_alpha[tau].c_str(), v.alpha, req, false);
                p.get(n_beta[tau].c_str(), v.beta, req, false);
                if (!p.get(n_uses[tau].c_str(), v.allowed_uses, false)) p.get("allowed_uses", v.allowed_uses, false);
                p.get(n_preload[tau].c_str(), v.preload, false, false);
                for (int at = 0; at < AT; ++at) p.get(n_toll[(size_t)tau * AT + at].c_str(), v.toll()[at], false, false);
                p.get(n_penalty[tau].c_str(), v.penalty, false, false);
                if (cell_type >= 2) v.penalty += 1.0 / 60.0;
                p.get("cycle_length", v.cycle, false, false);
                if (v.cycle >= 1) {
                    int g0 = -1, g1 = -1;
                    p.get("start_green_time", g0);
                    p.get("end_green_time", g1);
                    float green = g1 - g0;
                    if (green < 0) green = v.cycle;
                    v.red = std::max(1.0f, v.cycle - green);
                    p.get("red_time", v.red, false);
                    if (sat_in > 1000) v.sat = sat_in;
                }
                per[tau] = std::move(v);
            }
            double ref_vol = -1;
            p.get(n_ref.c_str(), ref_vol, false, false);
            p.get("tmc", tmc, false);
            int tmc_cid = -1, tmc_seq = -1;
            if (!tmc.empty()) { tmc_cid = 1; tmc_seq = 1; p.get("tmc_corridor_id", tmc_cid, false); p.get("tmc_road_sequence", tmc_seq, false); }
            out.link_from.push_back(from); out.link_to.push_back(to); out.link_tag.push_back(tag);
            out.link_id.push_back(id); out.geometry.push_back(geometry); out.link_type_code.push_back(type_code); out.vdf_code.push_back(vdf_code);
            out.tmc_code.push_back(tmc); out.tmc_corridor_name.push_back(tmc_name); out.tmc_corridor_id.push_back(tmc_cid);
            out.tmc_road_sequence.push_back(tmc_seq); out.link_type.push_back(link_type); out.link_vdf_type.push_back(lt.vdf_type);
            out.meso_link_id.push_back(meso >= 0 ? meso : -1);
            out.free_speed.push_back(free_speed); out.cutoff_speed.push_back((float)cutoff); out.dist_km.push_back(length_m / 1000.0);
            out.dist_mile.push_back(dist); out.fftt_min.push_back(length_m / 1609.0 / free_speed * 60); out.ref_volume.push_back(ref_vol);
            lp_out.push_back(std::move(per));
        }
        };
        auto append = [&](dtl_gmns& o, std::vector<std::vector<PerPeriod>>& lo) {
#define APP(f) g.f.insert(g.f.end(), std::make_move_iterator(o.f.begin()), std::make_move_iterator(o.f.end()))
            APP(link_from); APP(link_to); APP(link_tag); APP(link_id); APP(geometry); APP(link_type_code); APP(vdf_code); APP(tmc_code);
            APP(tmc_corridor_name); APP(tmc_corridor_id); APP(tmc_road_sequence); APP(link_type); APP(link_vdf_type); APP(meso_link_id);
            APP(free_speed); APP(cutoff_speed); APP(dist_km); APP(dist_mile); APP(fftt_min); APP(ref_volume);
#undef APP
            lp.insert(lp.end(), std::make_move_iterator(lo.begin()), std::make_move_iterator(lo.end()));
        };
        std::string text;
        Csv::read_file(g.dir + "/link.csv", text);
        const size_t h_end = std::min(text.find('\n'), text.size());
        const std::string header = text.substr(0, h_end);
        const char* body = text.data() + std::min(h_end + 1, text.size());
        const char* body_end = text.data() + text.size();
        {   // an empty line ends the reference's ReadRecord loop: nothing after it is read
            const char* q = body;
            while (q < body_end) {
                const char* nl = (const char*)memchr(q, '\n', (size_t)(body_end - q));
                if (nl == q) { body_end = q; break; }
                if (!nl) break;
                q = nl + 1;
            }
        }
        unsigned n_thr = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u);
        if (const char* env = getenv("DTL_IO_THREADS")) n_thr = (unsigned)std::max(1, atoi(env));
        if ((size_t)(body_end - body) < ((size_t)1 << 20)) n_thr = 1;
        bool redo_serial = n_thr == 1;
        if (n_thr > 1) {
            std::vector<const char*> cut(n_thr + 1, body_end);
            cut[0] = body;
            for (unsigned i = 1; i < n_thr; ++i) {
                const char* q = body + (size_t)(body_end - body) * i / n_thr;
                q = std::max(q, cut[i - 1]);
                const char* nl = (const char*)memchr(q, '\n', (size_t)(body_end - q));
                cut[i] = nl ? nl + 1 : body_end;
            }
            std::vector<dtl_gmns> outs(n_thr);
            std::vector<std::vector<std::vector<PerPeriod>>> lps(n_thr);
            std::vector<std::string> errs(n_thr);
            std::vector<char> unknown(n_thr, 0);
            std::vector<std::thread> pool;
            for (unsigned i = 0; i < n_thr; ++i)
                pool.emplace_back([&, i]() {
                    try {
                        Csv p;
                        p.open_text(cut[i], cut[i + 1], header, g.dir + "/link.csv");
                        std::map<int, LinkType> types = link_types;
                        bool unk = false;
                        parse_rows(p, outs[i], lps[i], types, false, unk);
                        unknown[i] = unk;
                    } catch (const StopError& e) { errs[i] = e.what(); }
                });
            for (auto& t : pool) t.join();
            for (unsigned i = 0; i < n_thr; ++i) {
                if (!errs[i].empty()) stop(errs[i]);
                redo_serial = redo_serial || unknown[i];
            }
            if (!redo_serial)
                for (unsigned i = 0; i < n_thr; ++i) append(outs[i], lps[i]);
        }
        if (redo_serial) {
            Csv p;
            p.open_text(body, body_end, header, g.dir + "/link.csv");
            dtl_gmns out;
            std::vector<std::vector<PerPeriod>> lo;
            bool unk = false;
            parse_rows(p, out, lo, link_types, true, unk);
            append(out, lo);
        }
    }
    // ---- link_qvdf.csv defaults, input.h:2153-2300 (no per-link overrides supported) ----
    if (std::ifstream(g.dir + "/link_qvdf.csv").good()) stop("link_qvdf.csv is not supported yet");
    for (auto& per : lp)
        for (int tau = 0; tau < T; ++tau)
            if (per[tau].qdf < 0.0) per[tau].qdf = 1.0 / std::max(0.01, per[tau].L);
    // ---- virtual connectors, main_api.cpp:534-570 + input.h:1859-1908 (created after the q-vdf defaults: qdf stays -1) ----
    for (int i = 0; i < n_phys; ++i) {
        const int zid = g.node_zone_org[i];
        if (zid < 1) continue;
        const int zs = zone_seq[zid];
        const int centroid = g.zone_centroid[zs];
        for (int dir = 0; dir < 2; ++dir) {
            std::vector<PerPeriod> per(T);
            for (int tau = 0; tau < T; ++tau) {
                PerPeriod v = default_period(tau);
                v.cap_lane = 99999; v.fftt = 0.0001; v.alpha = 0; v.beta = 0;
                per[tau] = std::move(v);
            }
            g.link_from.push_back(dir == 0 ? i : centroid);
            g.link_to.push_back(dir == 0 ? centroid : i);
            g.link_tag.push_back(dir == 0 ? -1 : zs);
            g.link_id.push_back("connector"); g.geometry.push_back(""); g.link_type_code.push_back(""); g.vdf_code.push_back("");
            g.tmc_code.push_back(""); g.tmc_corridor_name.push_back(""); g.tmc_corridor_id.push_back(-1); g.tmc_road_sequence.push_back(-1);
            g.link_type.push_back(-1); g.link_vdf_type.push_back(0); g.meso_link_id.push_back(-1);
            g.free_speed.push_back(100); g.cutoff_speed.push_back(100); g.dist_km.push_back(0); g.dist_mile.push_back(0);
            g.fftt_min.push_back(0.1); g.ref_volume.push_back(-1);
            lp.push_back(std::move(per));
        }
    }
    // ---- transpose into the [T][L] / [T][AT][L] image ----
    const size_t L = lp.size();
    const size_t TL = (size_t)T * L, TAL = (size_t)T * AT * L;
    for (auto* v : { &g.fftt, &g.cap, &g.alpha, &g.beta, &g.vf, &g.vc, &g.nlanes, &g.cap_lane, &g.qdf, &g.preload, &g.penalty, &g.Q_alpha,
                     &g.Q_beta, &g.Q_cd, &g.Q_n, &g.Q_cp, &g.Q_s, &g.t2, &g.L_h, &g.start_h, &g.end_h, &g.plf })
        v->assign(TL, 0.0);
    g.cycle_length.assign(TL, 0); g.red_time.assign(TL, 0); g.sat_flow.assign(TL, 0); g.vdf_type.assign(TL, 0);
    g.pce.assign(TAL, 0); g.occ.assign(TAL, 0); g.toll.assign(TAL, 0); g.lr_price.assign(TAL, 0); g.allowed.assign(TAL, 0);
    g.link_allowed_uses.resize(L);
    for (size_t l = 0; l < L; ++l)
        for (int tau = 0; tau < T; ++tau) {
            const PerPeriod& v = lp[l][tau];
            const size_t i = (size_t)tau * L + l;
            g.fftt[i] = v.fftt; g.cap[i] = v.cap; g.alpha[i] = v.alpha; g.beta[i] = v.beta; g.vf[i] = v.vf; g.vc[i] = v.vc;
            g.nlanes[i] = v.nlanes; g.cap_lane[i] = v.cap_lane; g.qdf[i] = v.qdf; g.preload[i] = v.preload; g.penalty[i] = v.penalty;
            g.Q_alpha[i] = v.Q_alpha; g.Q_beta[i] = v.Q_beta; g.Q_cd[i] = v.Q_cd; g.Q_n[i] = v.Q_n; g.Q_cp[i] = v.Q_cp; g.Q_s[i] = v.Q_s;
            g.t2[i] = v.t2; g.L_h[i] = v.L; g.start_h[i] = v.start_h; g.end_h[i] = v.end_h; g.plf[i] = v.plf;
            g.cycle_length[i] = v.cycle; g.red_time[i] = v.red; g.sat_flow[i] = v.sat; g.vdf_type[i] = v.vdf_type;
            if (tau == 0) g.link_allowed_uses[l] = v.allowed_uses;
            for (int at = 0; at < AT; ++at) {
                const size_t a = ((size_t)tau * AT + at) * L + l;
                g.pce[a] = v.pce()[at]; g.occ[a] = v.occ()[at]; g.toll[a] = v.toll()[at]; g.lr_price[a] = v.lr()[at];
                // CLink::AllowAgentType, DTA.h:1352-1367: empty or "all" admits everyone, else substring match
                g.allowed[a] = v.allowed_uses.empty() || v.allowed_uses == "all" || v.allowed_uses.find(g.agents[at].name) != std::string::npos;
            }
        }
}

void read_demand(dtl_gmns& g)
{
    const int Z = (int)g.zone_id.size(), AT = (int)g.agents.size(), T = (int)g.periods.size();
    if (std::ifstream(g.dir + "/subarea.csv").good()) {
        Csv s;
        if (s.open(g.dir + "/subarea.csv") && s.read_record()) stop("subarea.csv focusing is not supported");
    }
    std::map<int, int> zone_seq;
    for (int z = 0; z < Z; ++z) zone_seq[g.zone_id[z]] = z;
    std::map<std::string, int> period_no, agent_no;
    for (int t = 0; t < T; ++t) period_no[g.periods[t].name] = t;
    for (int a = 0; a < AT; ++a) agent_no[g.agents[a].name] = a;
    std::map<long long, double> od; // key ((o*Z+d)*AT+at)*T+tau -> od_volume (double accumulator, DTA.h column pool)
    g.origin_demand.assign(Z, 0.0f);
    float total = 0;
    Csv p; p.first_line_header = false;
    if (!p.open(g.dir + "/settings.csv")) stop("File settings.csv cannot be opened.");
    while (p.read_record_section()) { // input.h:899-1203
        if (p.section != "[demand_file_list]") continue;
        int seq = 1;
        if (!p.get("file_sequence_no", seq)) break;
        if (seq <= -1) continue;
        double scale = 1.0;
        std::string file, period, format = "null", agent;
        p.get("file_name", file);
        p.get("demand_period", period);
        p.get("format_type", format);
        p.get("scale_factor", scale, false);
        p.get("agent_type", agent);
        if (!period_no.count(period)) stop("demand period = " + period + " in section [demand_file_list] has not been defined.");
        const int tau = period_no[period];
        g.periods[tau].n_files++;
        if (format.find("null") != std::string::npos) stop("Please provide format_type in section [demand_file_list.]");
        if (!agent_no.count(agent)) stop("agent_type = " + agent + " in section [demand_file_list] cannot be found.");
        const int at = agent_no[agent];
        if (format.find("column") == std::string::npos) stop("demand format_type '" + format + "' is not supported (column only)");
        Csv d;
        if (!d.open(g.dir + "/" + file)) stop("File " + file + " cannot be opened.");
        int o_id = 0, d_id = 0; // declared outside the loop like the reference: a blank field keeps the previous row's value
        while (d.read_record()) {
            float v = 0;
            d.get("o_zone_id", o_id);
            d.get("d_zone_id", d_id);
            d.get("volume", v);
            if (o_id == d_id) continue;
            auto oi = zone_seq.find(o_id), di = zone_seq.find(d_id);
            if (oi == zone_seq.end() || di == zone_seq.end()) continue;
            v *= (scale * 1.0); // float *= (double loading factor * double local factor), input.h:1146
            od[(((long long)oi->second * Z + di->second) * AT + at) * T + tau] += v;
            total += v;
            g.origin_demand[oi->second] += v;
        }
    }
    for (auto& kv : od) {
        if (!(kv.second > 0)) continue;
        long long k = kv.first;
        const int tau = (int)(k % T); k /= T;
        const int at = (int)(k % AT); k /= AT;
        const int d = (int)(k % Z);
        const int o = (int)(k / Z);
        g.od_o.push_back(o); g.od_d.push_back(d); g.od_at.push_back(at); g.od_tau.push_back(tau); g.od_vol.push_back(kv.second);
    }
    g.prob.total_demand_volume = total;
}

void read_output_configuration(dtl_gmns& g)
{   // [output_file_configuration], input.h:1790-1810
    Csv p; p.first_line_header = false;
    if (!p.open(g.dir + "/settings.csv")) return;
    while (p.read_record_section()) {
        if (p.section != "[output_file_configuration]") continue;
        p.get("path_output", g.path_output, false, false);
        break;
    }
}

void bind_problem(dtl_gmns& g, int blocks)
{
    dtl_problem& q = g.prob;
    q.n_nodes = (int)g.node_id.size(); q.n_links = (int)g.link_from.size(); q.n_zones = (int)g.zone_id.size();
    q.n_agent_types = (int)g.agents.size(); q.n_periods = (int)g.periods.size();
    q.memory_blocks = std::min(std::max(1, blocks), 100); // main_api.cpp:519, MAX_MEMORY_BLOCKS DTA.h:20
    g.vot.clear();
    for (const AgentType& a : g.agents) g.vot.push_back(a.vot);
    q.link_from = g.link_from.data(); q.link_to = g.link_to.data(); q.link_tag = g.link_tag.data();
    q.node_zone = g.node_zone.data(); q.zone_centroid = g.zone_centroid.data();
    q.fftt = g.fftt.data(); q.cap = g.cap.data(); q.alpha = g.alpha.data(); q.beta = g.beta.data(); q.vf = g.vf.data(); q.vc = g.vc.data();
    q.nlanes = g.nlanes.data(); q.cap_lane = g.cap_lane.data(); q.qdf = g.qdf.data(); q.preload = g.preload.data();
    q.penalty = g.penalty.data(); q.Q_alpha = g.Q_alpha.data(); q.Q_beta = g.Q_beta.data(); q.Q_cd = g.Q_cd.data(); q.Q_n = g.Q_n.data();
    q.Q_cp = g.Q_cp.data(); q.Q_s = g.Q_s.data(); q.t2 = g.t2.data(); q.L_h = g.L_h.data(); q.start_h = g.start_h.data();
    q.end_h = g.end_h.data(); q.plf = g.plf.data(); q.cycle_length = g.cycle_length.data(); q.red_time = g.red_time.data();
    q.sat_flow = g.sat_flow.data(); q.vdf_type = g.vdf_type.data(); q.pce = g.pce.data(); q.occ = g.occ.data(); q.toll = g.toll.data();
    q.lr_price = g.lr_price.data(); q.allowed = g.allowed.data(); q.vot = g.vot.data();
    q.n_od = (int)g.od_o.size();
    q.od_o = g.od_o.data(); q.od_d = g.od_d.data(); q.od_at = g.od_at.data(); q.od_tau = g.od_tau.data(); q.od_vol = g.od_vol.data();
    q.origin_demand = g.origin_demand.data();
}

// g_load_supply_side_scenario_file, scenario_API.h:53-190: rows of type "sa" change the lanes of a link for the
// sensitivity-analysis stage; "dms" rows (scenario_API.h:165-188) mark a link as a message sign and the zone of its to-node as an
// information zone (dtl_gmns_dms -> dtl_sa_set_dms / dtl_sa_route_modification).
void read_scenario(dtl_gmns& g)
{
    Csv p;
    if (!p.open(g.dir + "/supply_side_scenario.csv")) return;
    std::map<int, int> node_seq;
    for (size_t i = 0; i < g.node_id.size(); ++i)
        if (!node_seq.count(g.node_id[i])) node_seq[g.node_id[i]] = (int)i;
    std::map<std::pair<int, int>, int> link_of; // m_to_node_2_link_seq_no_map: the last link between two nodes
    for (size_t l = 0; l < g.link_from.size(); ++l) link_of[{ g.link_from[l], g.link_to[l] }] = (int)l;
    while (p.read_record()) {
        int from_id = 0, to_id = 0;
        if (!p.get("from_node_id", from_id)) continue;
        if (!p.get("to_node_id", to_id)) continue;
        if (from_id == 0 && to_id == 0) continue;
        auto fi = node_seq.find(from_id), ti = node_seq.find(to_id);
        if (fi == node_seq.end() || ti == node_seq.end()) continue;
        auto li = link_of.find({ fi->second, ti->second });
        if (li == link_of.end()) continue;
        std::string period, type;
        p.get("demand_period", period, false);
        int tau = 0;
        for (size_t t = 0; t < g.periods.size(); ++t)
            if (g.periods[t].name == period) tau = (int)t;
        p.get("scenario_type", type, false);
        if (type == "sa") {
            float lanes = 0;
            p.get("lanes", lanes, false);
            g.sa_tau.push_back(tau); g.sa_link.push_back(li->second); g.sa_lanes.push_back(lanes);
            if (g.design_flag.empty()) g.design_flag.assign(g.periods.size() * g.link_from.size(), 0);
            g.design_flag[(size_t)tau * g.link_from.size() + li->second] = -1;
        } else if (type == "dms") {
            const int zid = g.node_zone_org[ti->second];
            if (zid < 1) stop("supply_side_scenario.csv: information zone has not been defined!"); // scenario_API.h:170-175
            int zone = -1;
            for (size_t z = 0; z < g.zone_id.size(); ++z)
                if (g.zone_id[z] == zid) zone = (int)z;                                   // g_zoneid_to_zone_seq_no_mapping, :179-182
            if (g.design_flag.empty()) g.design_flag.assign(g.periods.size() * g.link_from.size(), 0);
            g.design_flag[(size_t)tau * g.link_from.size() + li->second] = 2;
            g.dms_tau.push_back(tau); g.dms_link.push_back(li->second); g.dms_zone.push_back(zone);
        }
    }
}

// the "link" rows of measurement.csv, Assignment::Demand_ODME, ODME.h:100-178
void read_measurements(dtl_gmns& g)
{
    Csv p;
    if (!p.open(g.dir + "/measurement.csv")) return;
    std::map<int, int> node_seq;
    for (size_t i = 0; i < g.node_id.size(); ++i)
        if (!node_seq.count(g.node_id[i])) node_seq[g.node_id[i]] = (int)i;
    std::map<std::pair<int, int>, int> link_of; // m_to_node_2_link_seq_no_map: the last link between two nodes
    for (size_t l = 0; l < g.link_from.size(); ++l) link_of[{ g.link_from[l], g.link_to[l] }] = (int)l;
    const size_t L = g.link_from.size();
    g.obs_count.assign(g.periods.size() * L, 0.0);
    g.obs_ub.assign(g.periods.size() * L, 1); // CPeriod_VDF(): upper_bound_flag 1, VDF.h:60
    while (p.read_record()) {
        std::string type;
        p.get("measurement_type", type, false);
        if (type != "link") continue; // (the production / attraction branches of the reference sit inside the link branch: dead code)
        int from_id = 0, to_id = 0;
        if (!p.get("from_node_id", from_id)) continue;
        if (!p.get("to_node_id", to_id)) continue;
        auto fi = node_seq.find(from_id), ti = node_seq.find(to_id);
        if (fi == node_seq.end() || ti == node_seq.end()) continue;
        float count = -1;
        for (size_t tau = 0; tau < g.periods.size(); ++tau) {
            int ub = 0;
            if (g.periods[tau].n_files == 0) continue;
            p.get("count" + std::to_string(g.periods[tau].id), count, true, false);
            p.get("upper_bound_flag" + std::to_string(g.periods[tau].id), ub, true, false);
            auto li = link_of.find({ fi->second, ti->second });
            if (li == link_of.end()) continue;
            g.ms_tau.push_back((int)tau); g.ms_link.push_back(li->second); g.ms_count.push_back(count); g.ms_ub.push_back(ub);
            const size_t a = tau * L + li->second;                 // ODME.h:156-175
            if (g.obs_count[a] >= 1) {
                if (ub == 0) { g.obs_count[a] = count; g.obs_ub[a] = ub; }
            } else {
                g.obs_count[a] = count; g.obs_ub[a] = ub;
            }
        }
    }
}

} // namespace

extern "C" {

dtl_gmns* dtl_gmns_read(const char* dir, int number_of_memory_blocks)
{
    dtl_gmns* g = new dtl_gmns();
    g->dir = dir && *dir ? dir : ".";
    const bool trace = getenv("DTL_IO_TRACE") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    auto phase = [&](const char* what) {
        const auto now = std::chrono::steady_clock::now();
        if (trace) fprintf(stderr, "gmns_read: %s %.3f s\n", what, std::chrono::duration<double>(now - t0).count());
        t0 = now;
    };
    try {
        read_settings(*g);
        phase("settings.csv");
        read_network(*g);
        phase("node.csv + link.csv + connectors");
        read_demand(*g);
        phase("demand files");
        read_output_configuration(*g);
        read_scenario(*g);
        read_measurements(*g);
        bind_problem(*g, number_of_memory_blocks);
        phase("problem image");
    } catch (const StopError& e) {
        g_io_err = e.what();
        delete g;
        return nullptr;
    }
    return g;
}

const dtl_problem* dtl_gmns_problem(const dtl_gmns* g) { return g ? &g->prob : nullptr; }
void dtl_gmns_free(dtl_gmns* g) { delete g; }
const char* dtl_gmns_last_error(void) { return g_io_err.c_str(); }

int dtl_gmns_scenario(const dtl_gmns* g, int* rt_info, int* sa_tau, int* sa_link, double* sa_lanes, int capacity)
{
    if (!g) return -1;
    if (rt_info)
        for (size_t at = 0; at < g->agents.size(); ++at) rt_info[at] = g->agents[at].rt_info;
    const int n = (int)g->sa_link.size();
    for (int i = 0; i < std::min(n, capacity); ++i) {
        if (sa_tau) sa_tau[i] = g->sa_tau[i];
        if (sa_link) sa_link[i] = g->sa_link[i];
        if (sa_lanes) sa_lanes[i] = g->sa_lanes[i];
    }
    return n;
}

int dtl_gmns_dms(const dtl_gmns* g, int* dms_tau, int* dms_link, int* dms_zone, int capacity)
{
    if (!g) return -1;
    const int n = (int)g->dms_link.size();
    for (int i = 0; i < std::min(n, capacity); ++i) {
        if (dms_tau) dms_tau[i] = g->dms_tau[i];
        if (dms_link) dms_link[i] = g->dms_link[i];
        if (dms_zone) dms_zone[i] = g->dms_zone[i];
    }
    return n;
}

int dtl_gmns_measurements(const dtl_gmns* g, float* at_pce, float* at_occ, int* ms_tau, int* ms_link, float* ms_count, int* ms_ub, int capacity)
{
    if (!g) return -1;
    for (size_t at = 0; at < g->agents.size(); ++at) {
        if (at_pce) at_pce[at] = (float)g->agents[at].pce; // float PCE_ratio = g_AgentTypeVector[at].PCE, assignment.h:308
        if (at_occ) at_occ[at] = (float)g->agents[at].occ;
    }
    const int n = (int)g->ms_link.size();
    for (int i = 0; i < std::min(n, capacity); ++i) {
        if (ms_tau) ms_tau[i] = g->ms_tau[i];
        if (ms_link) ms_link[i] = g->ms_link[i];
        if (ms_count) ms_count[i] = g->ms_count[i];
        if (ms_ub) ms_ub[i] = g->ms_ub[i];
    }
    return n;
}

int dtl_gmns_ids(const dtl_gmns* g, int* node_id, int* zone_id)
{
    if (!g) return -1;
    if (node_id) memcpy(node_id, g->node_id.data(), sizeof(int) * g->node_id.size());
    if (zone_id) memcpy(zone_id, g->zone_id.data(), sizeof(int) * g->zone_id.size());
    return 0;
}

// flash_dta.cpp:171-242: number_of_iterations | column_generation_iterations; measurement.csv with a 'link' row
// switches on 5 column-updating + 100 ODME iterations; supply_side_scenario.csv with a row switches on 20 SA iterations.
int dtl_gmns_assignment_settings(const char* dir_c, int out[6])
{
    const std::string dir = dir_c && *dir_c ? dir_c : ".";
    int cg = 0, cu = 0, odme = 0, sa = 0, sim = 0, blocks = 4;
    try {
        Csv p; p.first_line_header = false;
        if (!p.open(dir + "/settings.csv")) { g_io_err = "File settings.csv cannot be opened."; return -1; }
        if (p.read_record_section()) { // the reference takes the first record whatever the section (stray ';', flash_dta.cpp:175)
            p.get("number_of_iterations", cg, false);
            if (cg == 0) p.get("column_generation_iterations", cg, true, true);
            {   // CheckMeasurementFileExist, flash_dta.cpp:66-100
                Csv m;
                int link_rows = 0;
                if (m.open(dir + "/measurement.csv"))
                    while (m.read_record()) {
                        std::string t;
                        int a, b;
                        m.get("measurement_type", t);
                        if (t == "link" && m.get("from_node_id", a) && m.get("to_node_id", b)) ++link_rows;
                    }
                if (link_rows >= 1) { cu = 5; odme = 100; }
            }
            {   // CheckSupplySideScenarioFileExist, flash_dta.cpp:103-138
                Csv s;
                int rows = 0;
                if (s.open(dir + "/supply_side_scenario.csv"))
                    while (s.read_record()) {
                        int a = 0, b = 0;
                        s.get("from_node_id", a);
                        s.get("to_node_id", b);
                        if (a != 0 || b != 0) ++rows;
                    }
                if (rows) sa = 20;
            }
            p.get("simulation_iterations", sim, false, false);
            p.get("number_of_memory_blocks", blocks, false, false);
        }
    } catch (const StopError& e) {
        g_io_err = e.what();
        return -1;
    }
    out[0] = cg; out[1] = cu; out[2] = odme; out[3] = sa; out[4] = sim; out[5] = blocks;
    return 0;
}

// printf into a growing string (the writers format rows in parallel and write them in order)
static void appendf(std::string& out, const char* fmt, ...)
{
    char b[512];
    va_list ap;
    va_start(ap, fmt);
    int n = vsnprintf(b, sizeof b, fmt, ap);
    va_end(ap);
    if (n < 0) return;
    if ((size_t)n < sizeof b) { out.append(b, (size_t)n); return; }
    std::vector<char> big((size_t)n + 1);
    va_start(ap, fmt);
    vsnprintf(big.data(), big.size(), fmt, ap);
    va_end(ap);
    out.append(big.data(), (size_t)n);
}

// -------------------------------------------------------------------------------------------------
int dtl_gmns_write_outputs(const dtl_gmns* gp, dtl_engine* e, const char* dir_c)
{
    if (!gp || !e) return -1;
    const dtl_gmns& g = *gp;
    const std::string dir = dir_c && *dir_c ? dir_c : ".";
    const bool trace = getenv("DTL_IO_TRACE") != nullptr;
    auto t_phase = std::chrono::steady_clock::now();
    auto phase = [&](const char* what) {
        const auto now = std::chrono::steady_clock::now();
        if (trace) fprintf(stderr, "write_outputs: %s %.3f s\n", what, std::chrono::duration<double>(now - t_phase).count());
        t_phase = now;
    };
    const int L = g.prob.n_links, T = g.prob.n_periods, AT = g.prob.n_agent_types, Z = g.prob.n_zones;
    std::vector<std::vector<double>> tt(T), pce(T), person(T), lv(T), doc(T), voc(T), P(T), vt2(T), qs(T), sbpr(T), qmu(T), qg(T), t0(T),
        t3(T), sev(T);
    std::vector<std::vector<std::vector<double>>> person_at(T, std::vector<std::vector<double>>(AT));
    std::vector<std::vector<float>> speed(T);
    const int first_slot = 6 * 60 / 5, n_slots = (20 * 60 - 6 * 60) / 5 + 3;
    for (int tau = 0; tau < T; ++tau) {
        for (auto* v : { &tt[tau], &pce[tau], &person[tau], &lv[tau], &doc[tau], &voc[tau], &P[tau], &vt2[tau], &qs[tau], &sbpr[tau],
                         &qmu[tau], &qg[tau], &t0[tau], &t3[tau], &sev[tau] })
            v->assign(L, 0.0);
        if (dtl_get_link_state(e, tau, tt[tau].data(), pce[tau].data(), person[tau].data(), lv[tau].data(), doc[tau].data(),
                               voc[tau].data(), P[tau].data(), vt2[tau].data()))
            return -1;
        if (dtl_get_link_detail(e, tau, qs[tau].data(), sbpr[tau].data(), qmu[tau].data(), qg[tau].data(), t0[tau].data(), t3[tau].data(),
                                sev[tau].data()))
            return -1;
        for (int at = 0; at < AT; ++at) {
            person_at[tau][at].assign(L, 0.0);
            if (dtl_get_person_volume_at(e, tau, at, person_at[tau][at].data())) return -1;
        }
        speed[tau].assign((size_t)L * n_slots, 0.0f);
        if (dtl_get_model_speed(e, tau, first_slot, n_slots, speed[tau].data())) return -1;
    }
    phase("link state read-back");
    // ---- link_performance.csv, output.h:533-779 ----
    FILE* f = fopen((dir + "/link_performance.csv").c_str(), "w");
    if (!f) { g_io_err = "File link_performance.csv cannot be opened."; return -1; }
    fprintf(f, "link_id,vdf_type,from_node_id,to_node_id,lanes,distance_km,distance_mile,FFTT,meso_link_id,meso_link_incoming_volume,tmc,"
               "tmc_corridor_name,tmc_corridor_id,tmc_road_order,tmc_road_sequence,subarea_id,link_type,link_type_code,vdf_code,time_period,"
               "volume,ref_volume,ref_diff,odme_volume,obs_count,upper_bound_type,obs_count_dev,");
    fprintf(f, "preload_volume,person_volume,travel_time,speed,speed_ratio,VOC,DOC,capacity,queue,total_simu_waiting_time_in_min,"
               "avg_simu_waiting_time_in_min,qdf,lanes,D_per_hour_per_lane,QVDF_cd,QVDF_n,P,severe_congestion_duration_in_h,vf,"
               "v_congestion_cutoff,QVDF_cp,QVDF_s,QVDF_v,vt2,VMT,VHT,PMT,PHT,PDT_vf,PDT_vc,geometry,");
    for (int at = 0; at < AT; ++at) fprintf(f, "person_vol_%s,", g.agents[at].name.c_str());
    for (int t = 6 * 60; t < 20 * 60; t += 15) fprintf(f, "v%02d:%02d,", t / 60, t % 60);
    fprintf(f, "scenario_code,volume_before,volume_after,volume_diff,speed_before,speed_after,speed_diff,DoC_before,DoC_after,Doc_diff,"
               "P_before,P_after,P_diff,");
    fprintf(f, "notes\n");
    auto format_link = [&](int i, std::string& out) {
        if (g.link_type[i] == -1) return; // virtual connectors
        float total_vehicle_volume = 0;
        int scenario_code_count = 0;
        for (int tau = 0; tau < T; ++tau) {                           // output.h:587-598
            if (!g.before_volume.empty()) total_vehicle_volume += g.before_volume[(size_t)tau * L + i];
            total_vehicle_volume += pce[tau][i] + g.preload[(size_t)tau * L + i];
            if (!g.design_flag.empty() && g.design_flag[(size_t)tau * L + i] == -1) scenario_code_count += 2;
        }
        if (scenario_code_count == 0 && total_vehicle_volume < 0.001) return;
        for (int tau = 0; tau < T; ++tau) {
            if (g.periods[tau].n_files == 0) continue;
            const size_t k = (size_t)tau * L + i;
            const double avg_tt = tt[tau][i];
            float spd = g.free_speed[i];
            if (avg_tt > 0.001f) spd = g.free_speed[i] / std::max(0.000001, avg_tt) * g.fftt[k];
            const float speed_ratio = spd / std::max(1.0, g.free_speed[i]);
            const float vehicle_volume = pce[tau][i] + g.preload[k];
            float ref_diff = 0;
            if (g.ref_volume[i] > 1) ref_diff = vehicle_volume - g.ref_volume[i];
            const float person_volume = person[tau][i] + g.preload[k];
            const float preload = g.preload[k];
            const float VMT = vehicle_volume * g.dist_mile[i];
            const float VHT = vehicle_volume * avg_tt / 60.0;
            const float PMT = person_volume * g.dist_mile[i];
            const float PHT = person_volume * avg_tt / 60.0;
            const float PDT_vf = person_volume * (avg_tt - g.fftt[k]) / 60.0;
            const float PDT_vc = std::max(0.0, person_volume * (avg_tt - g.fftt[k] * g.free_speed[i] / std::max(0.001f, (float)g.cutoff_speed[i])) / 60.0);
            const double lane_based_D = std::max(0.0, lv[tau][i]) * g.qdf[k] / std::max(0.000001, g.nlanes[k]);
            appendf(out, "%s,%s,%d,%d,%f,%f,%f,%f,%d,%d,", g.link_id[i].c_str(), g.link_vdf_type[i] == 1 ? "bpr" : "qvdf",
                    g.node_id[g.link_from[i]], g.node_id[g.link_to[i]], g.nlanes[k], g.dist_km[i], g.dist_mile[i], g.fftt_min[i],
                    g.meso_link_id[i], 0);
            appendf(out, "%s,%s,%d,%d,%d,%d,%d,%s,%s,%s,", g.tmc_code[i].c_str(), g.tmc_corridor_name[i].c_str(), g.tmc_corridor_id[i], 0,
                    g.tmc_road_sequence[i], -1, g.link_type[i], g.link_type_code[i].c_str(), g.vdf_code[i].c_str(),
                    g.periods[tau].time_period.c_str());
            appendf(out, "%.3f,%.3f,%.3f,", vehicle_volume, g.ref_volume[i], ref_diff);
            if (g.odme_ran && g.obs_count[k] >= 1) {                   // output.h:682-692
                const double vbef = g.before_volume.empty() ? (double)vehicle_volume : g.before_volume[k];
                appendf(out, "%.3f,%.3f,%d,%.3f,", vbef, g.obs_count[k], g.obs_ub[k], g.obs_count[k] - vbef);
            } else {
                appendf(out, ",,,,");
            }
            appendf(out, "%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,", preload, person_volume, avg_tt, spd, speed_ratio, voc[tau][i],
                    doc[tau][i], g.cap_lane[k], 0.0, 0.0);
            appendf(out, "%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,", 0.0 / std::max(1.0f, vehicle_volume), g.qdf[k], g.nlanes[k],
                    lane_based_D, g.Q_cd[k], g.Q_n[k], P[tau][i], sev[tau][i], g.vf[k], g.vc[k]);
            appendf(out, "%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,%.3f,", g.Q_cp[k], g.Q_s[k], qs[tau][i], vt2[tau][i], VMT, VHT, PMT, PHT,
                    PDT_vf, PDT_vc);
            appendf(out, "\"%s\",", g.geometry[i].c_str());
            for (int at = 0; at < AT; ++at) appendf(out, "%.3f,", person_at[tau][at][i]);
            for (int t = 6 * 60; t < 20 * 60; t += 15) { // CLink::get_model_15_min_speed, DTA.h:1194-1214
                float sum = 0;
                int cnt = 0;
                for (int q = 0; q < 3; ++q) {
                    // model_speed[] is ONE array per link (DTA.h:1179): every demand period's VDF call writes the slots of its
                    // own time window (VDF.h:313), later periods last, and every row of the link prints the merged profile;
                    // slots outside all windows keep the reader's initial value, the free speed (input.h:3490-3496)
                    const int slot = t / 5 + q - first_slot, abs_slot = t / 5 + q;
                    float v = (float)g.free_speed[i];
                    for (int t2 = 0; t2 < T; ++t2) {
                        const size_t k2 = (size_t)t2 * L + i;
                        const int s0 = (int)(g.start_h[k2] * 60) / 5, s1 = (int)(g.end_h[k2] * 60) / 5;
                        if (abs_slot >= s0 && abs_slot <= s1) v = speed[t2][(size_t)i * n_slots + slot];
                    }
                    if (v >= 1) { sum += v; ++cnt; }
                }
                appendf(out, "%.3f,", sum / std::max(1, cnt));
            }
            appendf(out, "%s,", g.design_flag.empty() ? "" : g.design_flag[k] == -1 ? "sa" : g.design_flag[k] == 2 ? "dms" : ""); // scenario_code, scenario_API.h:160,168
            // base case / after the sensitivity-analysis stage (g_record_corridor_performance_summary, output.h:52-98): the
            // "before" record is taken between the assignment and the stage (network_assignment below), the "after" record is
            // the state being written.  Without the stage both are the final state and the differences are zero.
            const double va = vehicle_volume, sa = spd, da = doc[tau][i], pa = P[tau][i];
            const bool rec = !g.before_volume.empty();
            const double vb = rec ? g.before_volume[k] : va, sb = rec ? g.before_speed[k] : sa, db = rec ? g.before_doc[k] : da,
                         pb = rec ? g.before_P[k] : pa;
            appendf(out, "%.3f,%.3f,%.3f,", vb, va, va - vb);
            appendf(out, "%.3f,%.3f,%.3f,", sb, sa, sa - sb);
            appendf(out, "%.3f,%.3f,%.3f,", db, da, da - db);
            appendf(out, "%.3f,%.3f,%.3f,", pb, pa, pa - pb);
            appendf(out, "period-based\n");
        }
    };
    {
        const size_t chunk = 1024;
        const int n_threads = (int)std::max<size_t>(1, std::min<size_t>({ (size_t)32, (size_t)std::max(1u, std::thread::hardware_concurrency()),
                                                                           ((size_t)L + chunk - 1) / chunk }));
        std::vector<std::string> bufs(n_threads);
        for (size_t base = 0; base < (size_t)L; base += chunk * n_threads) {
            std::vector<std::thread> pool;
            for (int t = 0; t < n_threads; ++t) {
                const size_t r0 = base + chunk * t, r1 = std::min((size_t)L, r0 + chunk);
                bufs[t].clear();
                if (r0 >= r1) continue;
                pool.emplace_back([&, t, r0, r1] {
                    for (size_t r = r0; r < r1; ++r) format_link((int)r, bufs[t]);
                });
            }
            for (auto& th : pool) th.join();
            for (int t = 0; t < n_threads; ++t)
                if (!bufs[t].empty()) fwrite(bufs[t].data(), 1, bufs[t].size(), f);
        }
    }
    fclose(f);
    phase("link_performance.csv");
    // ---- route_assignment.csv, output.h:814-1456 ----
    f = fopen((dir + "/route_assignment.csv").c_str(), "w");
    if (!f) { g_io_err = "File route_assignment.csv cannot be opened."; return -1; }
    if (g.path_output == 0) { fclose(f); return 0; }
    fprintf(f, "first_column,path_no,o_zone_id,d_zone_id,od_pair,o_sindex,d_sindex,within_OD_path_no,path_id,activity_zone_sequence,"
               "activity_agent_type_sequence,information_type,agent_type,demand_period,volume,simu_volume,subarea_flag,OD_impact_flag,"
               "at_OD_impact_flag,");
    fprintf(f, "path_impact_flag,toll,#_of_nodes,travel_time,VDF_travel_time,VDF_travel_time_without_access_link,distance,distance_km,"
               "distance_mile,node_sequence,link_sequence, ");
    for (int it = 0; it <= g.sa_iterations; ++it) fprintf(f, "SA_TT_%d,", it);
    for (int it = 0; it <= g.sa_iterations; ++it) fprintf(f, it == 0 ? "SA_Vol_%d," : "SA_Delta_Vol_%d,", it);
    fprintf(f, "geometry,");
    fprintf(f, "link_type_name_sequence,link_code_sequence,link_link_distance_VDF_sequence,link_FFTT_sequence,");
    fprintf(f, "\n");
    const long long nc = dtl_num_columns(e), nl = dtl_num_column_links(e);
    if (nc < 0 || nl < 0) { fclose(f); return -1; }
    std::vector<int> od_index(std::max<long long>(nc, 1)), node_sum(od_index.size()), seq(od_index.size()), n_links(od_index.size()),
        n_nodes(od_index.size()), links(std::max<long long>(nl, 1));
    std::vector<double> volume(od_index.size()), cost(od_index.size()), time(od_index.size());
    if (dtl_get_columns(e, od_index.data(), node_sum.data(), seq.data(), n_links.data(), n_nodes.data(), volume.data(), cost.data(),
                        time.data(), links.data())) { fclose(f); return -1; }
    // columns arrive grouped by OD cell in (o, d, at, tau) order; the reference nests o, d, tau, at
    std::vector<long long> first_link(nc + 1, 0);
    for (long long c = 0; c < nc; ++c) first_link[c + 1] = first_link[c] + n_links[c];
    std::vector<long long> order(nc);
    for (long long c = 0; c < nc; ++c) order[c] = c;
    std::stable_sort(order.begin(), order.end(), [&](long long a, long long b) {
        const int ia = od_index[a], ib = od_index[b];
        if (g.od_o[ia] != g.od_o[ib]) return g.od_o[ia] < g.od_o[ib];
        if (g.od_d[ia] != g.od_d[ib]) return g.od_d[ia] < g.od_d[ib];
        if (g.od_tau[ia] != g.od_tau[ib]) return g.od_tau[ia] < g.od_tau[ib];
        return g.od_at[ia] < g.od_at[ib];
    });
    (void)Z;
    phase("column read-back and ordering");
    // Rows are formatted in parallel and written in order: the file is ~2.5 KB per column (node, link and coordinate
    // sequences), gigabytes on a regional network, and one fprintf stream made it the longest phase of the whole run.
    // Per-node / per-link text is formatted once; the row prefix keeps the reference's printf format (output.h:1248-1445).
    // message signs ("dms" rows): the route modification's impacted_path_flag per column and information_type per cell -- a cell with
    // information_type 1 is printed from its first physical node on (no virtual first link is stripped) and without its columns
    // below 0.1 (output.h:1219-1226)
    std::vector<unsigned char> route_impact(od_index.size(), 0), info_type(od_index.size(), 0);
    const bool with_dms = g.sa_stage && !g.dms_link.empty();
    if (with_dms && dtl_sa_get_route_flags(e, route_impact.data(), info_type.data())) { fclose(f); return -1; }
    std::vector<long long> kept; // columns that are written, in output order: more than two physical nodes (output.h:1239)
    kept.reserve(order.size());
    for (long long oc : order) {
        if (info_type[oc] && volume[oc] < 0.1) continue;
        if (n_nodes[oc] - (info_type[oc] ? 1 : 2) > 2) kept.push_back(oc);
    }
    const int N = g.prob.n_nodes;
    std::vector<std::string> node_seq(N), node_xy(N), link_seq(L);
    {
        char b[160];
        for (int n = 0; n < N; ++n) {
            snprintf(b, sizeof b, "%d;", g.node_id[n]);
            node_seq[n] = b;
            snprintf(b, sizeof b, "%f %f", g.node_x[n], g.node_y[n]);
            node_xy[n] = b;
        }
        for (int l = 0; l < L; ++l) link_seq[l] = g.link_id[l] + ";";
    }
    std::string sa_cols;
    {
        char b[64];
        for (int it = 0; it <= g.sa_iterations; ++it) { snprintf(b, sizeof b, "%f,", -1.0); sa_cols += b; }
        for (int it = 0; it <= g.sa_iterations; ++it) { snprintf(b, sizeof b, "%f,", 0.0); sa_cols += b; }
    }
    // Sensitivity-analysis records (only when the stage ran): per column the path travel time and volume after every SA column-pool
    // iteration (path_time_per_iteration_SA_map / path_volume_per_iteration_SA_map, assignment.h:741, 896-900), the cell's
    // at_od_impacted flag (assignment.h:699-713), and the flags g_classification_in_column_pool (assignment.h:930-1010) and
    // g_column_pool_route_modification (assignment.h:1176-1180) derive from the links with a lane change.
    int n_hist = 0;
    std::vector<unsigned char> at_impact(od_index.size(), 0), path_impact(od_index.size(), 0);
    std::vector<double> hist_time, hist_volume;
    std::map<std::tuple<int, int, int>, int> od_impact_regular; // (o, d, tau) -> a cell of an agent type without information is impacted
    if (g.sa_stage) {
        if (dtl_sa_get_columns(e, at_impact.data(), &n_hist, nullptr, nullptr)) { fclose(f); return -1; }
        hist_time.assign((size_t)std::max(1, n_hist) * od_index.size(), 0.0);
        hist_volume.assign((size_t)std::max(1, n_hist) * od_index.size(), 0.0);
        if (dtl_sa_get_columns(e, nullptr, nullptr, hist_time.data(), hist_volume.data())) { fclose(f); return -1; }
        if (!g.design_flag.empty())
            for (long long c = 0; c < nc; ++c) {
                const int cell = od_index[c], at = g.od_at[cell], tau = g.od_tau[cell];
                bool passes = false;
                for (long long q = first_link[c]; q < first_link[c + 1] && !passes; ++q) passes = g.design_flag[(size_t)tau * L + links[q]] == -1;
                if (!passes) continue;
                if (g.agents[at].rt_info == 0) {
                    path_impact[c] = 1;
                    od_impact_regular[std::make_tuple(g.od_o[cell], g.od_d[cell], tau)] = 1;
                } else if (g.agents[at].rt_info == 2 && !(volume[c] < 0.00001)) {
                    path_impact[c] = 1;
                }
            }
        if (with_dms) // the flags as g_column_pool_route_modification set them (on the volumes of that moment), incl. 3 = diverted new column
            for (long long c = 0; c < nc; ++c)
                if (g.agents[g.od_at[od_index[c]]].rt_info == 2) path_impact[c] = route_impact[c];
    }
    auto format_row = [&](long long oc, int count, std::string& out) {
        const int cell = od_index[oc];
        const int o = g.od_o[cell], d = g.od_d[cell], at = g.od_at[cell], tau = g.od_tau[cell];
        const int* pl = links.data() + first_link[oc];
        const int m_links = n_links[oc], m_nodes = n_nodes[oc];
        float path_toll = 0, path_tt = 0, path_tt_no_access = 0;
        double dist = 0, dist_km = 0, dist_ml = 0;
        for (int q = 0; q < m_links; ++q) {
            const int l = pl[q];
            if (g.link_type[l] >= 0) {
                path_toll += g.toll[((size_t)tau * AT + at) * L + l];
                dist += std::max(0.00001, g.dist_mile[l]); dist_km += g.dist_km[l]; dist_ml += g.dist_mile[l];
                const float ltt = tt[tau][l];
                path_tt += ltt;
                if (g.link_type[l] != 1000) path_tt_no_access += ltt;
            }
        }
        const double vol = std::max(0.001, volume[oc]);
        char b[1024];
        const int nb = snprintf(b, sizeof b, ",%d,%d,%d,%d,%d,%d->%d,%d,%d,%s,%s,%d,%s,%s,%.4f,%d,%d,%d,%d,%d,%.1f,%d,%.1f,%.4f,%.4f,%.4f,%.4f,%.4f,",
                                count, g.zone_id[o], g.zone_id[d], o, d, g.zone_id[o], g.zone_id[d], seq[oc], seq[oc] + 1, "", "", (int)info_type[oc],
                                g.agents[at].name.c_str(), g.periods[tau].name.c_str(), vol, 0, 0,
                                g.sa_stage && od_impact_regular.count(std::make_tuple(o, d, tau)) ? 1 : 0, (int)at_impact[oc], (int)path_impact[oc],
                                path_toll, m_nodes - (info_type[oc] ? 1 : 2),
                                (double)path_tt, path_tt, path_tt_no_access, dist, dist_km, dist_ml);
        out.append(b, std::min<size_t>(nb > 0 ? nb : 0, sizeof b - 1));
        // nodes: origin centroid, then the head of every link; the virtual first/last connector is stripped (the first one is kept
        // for a cell with information_type 1: virtual_first_link_delta = 0, output.h:1222-1226)
        const int first_delta = info_type[oc] ? 0 : 1;
        if (!first_delta && m_links > 0) out += node_seq[g.link_from[pl[0]]];
        for (int q = 0; q + 1 < m_links; ++q) out += node_seq[g.link_to[pl[q]]];
        out += ',';
        for (int q = first_delta; q + 1 < m_links; ++q) out += link_seq[pl[q]];
        out += ',';
        if (n_hist == 0) out += sa_cols;
        else {                                                        // output.h:1367-1390
            char hb[64];
            for (int it = 0; it <= g.sa_iterations; ++it) {
                snprintf(hb, sizeof hb, "%f,", it < n_hist ? hist_time[(size_t)it * od_index.size() + oc] : -1.0);
                out += hb;
            }
            double prev = 0;
            for (int it = 0; it <= g.sa_iterations; ++it) {
                const double v = it < n_hist ? hist_volume[(size_t)it * od_index.size() + oc] : 0.0;
                snprintf(hb, sizeof hb, "%f,", v - prev);
                out += hb;
                prev = v;
            }
        }
        if (volume[oc] >= 5 || count < 3000) {
            out += "\"LINESTRING (";
            if (!first_delta && m_links > 0) { out += node_xy[g.link_from[pl[0]]]; if (m_links > 1) out += ", "; }
            for (int q = 0; q + 1 < m_links; ++q) {
                out += node_xy[g.link_to[pl[q]]];
                if (q + 2 < m_links) out += ", ";
            }
            out += ")\"";
        }
        if (volume[oc] >= 5) out += ",,,,";
        out += '\n';
    };
    const size_t n_rows = kept.size();
    const size_t chunk = 2048;
    const int n_threads = (int)std::max<size_t>(1, std::min<size_t>({ (size_t)32, (size_t)std::max(1u, std::thread::hardware_concurrency()),
                                                                       (n_rows + chunk - 1) / chunk }));
    std::vector<std::string> bufs(n_threads);
    bool write_failed = false;
    for (size_t base = 0; base < n_rows && !write_failed; base += chunk * n_threads) {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t) {
            const size_t r0 = base + chunk * t, r1 = std::min(n_rows, r0 + chunk);
            bufs[t].clear();
            if (r0 >= r1) continue;
            pool.emplace_back([&, t, r0, r1] {
                for (size_t r = r0; r < r1; ++r) format_row(kept[r], (int)r + 1, bufs[t]);
            });
        }
        for (auto& th : pool) th.join();
        for (int t = 0; t < n_threads; ++t)
            if (!bufs[t].empty() && fwrite(bufs[t].data(), 1, bufs[t].size(), f) != bufs[t].size()) write_failed = true;
    }
    if (write_failed) { fclose(f); g_io_err = "File route_assignment.csv cannot be written."; return -1; }
    phase("route_assignment.csv rows");
    fclose(f);
    return 0;
}

// -------------------------------------------------------------------------------------------------
double network_assignment(int assignment_mode, int column_generation_iterations, int column_updating_iterations, int ODME_iterations,
                          int sensitivity_analysis_iterations, int simulation_iterations, int number_of_memory_blocks)
{
    (void)assignment_mode;
    auto fail = [](const std::string& m) {
        fprintf(stderr, "%s\nProgram stops.\n", m.c_str());
        std::ofstream log("log.txt", std::ios::app);
        log << m << "\nProgram stops." << std::endl;
        exit(2);
    };
    if (simulation_iterations > 0) fail("the mesoscopic simulator is outside the accelerated path (simulation_iterations must be 0)");
    using clk = std::chrono::steady_clock;
    const auto t_start = clk::now();
    auto since = [](clk::time_point t0) { return std::chrono::duration<double>(clk::now() - t0).count(); };
    dtl_gmns* g = dtl_gmns_read(".", number_of_memory_blocks);
    if (!g) fail(g_io_err);
    const double s_read = since(t_start);
    g->sa_iterations = std::max(0, sensitivity_analysis_iterations);
    const int S = g->sa_iterations;
    // stage II (main_api.cpp:771-1042) always runs in the reference; it only changes anything when an agent type has real-time
    // information, a supply-side scenario changes lanes, or SA iterations are asked for
    bool any_rt = false;
    for (const AgentType& a : g->agents) any_rt = any_rt || a.rt_info == 1;
    const bool sa_stage = any_rt || !g->sa_link.empty() || !g->dms_link.empty() || S > 0;
    dtl_options opt;
    dtl_default_options(&opt);
    opt.max_columns_per_od = std::max(4, column_generation_iterations + 4) + (sa_stage ? S + 1 : 0) + (g->dms_link.empty() ? 0 : 8);
    const auto t_create = clk::now();
    dtl_engine* e = dtl_create(dtl_gmns_problem(g), &opt);
    if (!e) fail(dtl_last_error());
    const double s_create = since(t_create);
    const auto t_loop = clk::now();
    std::ofstream summary("final_summary.csv");
    printf("DTALite (B200): %d nodes, %d links, %d zones, %d agent types, %d OD cells, %d shortest-path trees per iteration\n",
           g->prob.n_nodes, g->prob.n_links, g->prob.n_zones, g->prob.n_agent_types, g->prob.n_od, dtl_num_tasks(e));
    summary << "Step 1: read network node.csv, link.csv, zone.csv " << std::endl;
    summary << ",# of nodes = ," << g->prob.n_nodes << std::endl;
    summary << ",# of links =," << g->prob.n_links << std::endl;
    summary << ",# of zones =," << g->prob.n_zones << std::endl;
    std::vector<double> iter_rows((size_t)std::max(1, column_generation_iterations) * 5);
    if (dtl_iterations(e, 0, column_generation_iterations, iter_rows.data())) fail(dtl_last_error()); // main_api.cpp:589-728, pipelined
    for (int k = 0; k < column_generation_iterations; ++k) {
        const double* row = iter_rows.data() + (size_t)k * 5;
        if (k == 0) summary << "Iteration, CPU Running Time, # of agents, Avg Travel Time(min),  Avg UE gap %" << std::endl;
        summary << k << "," << 0.0 << "," << g->prob.total_demand_volume << "," << row[3] << "," << row[2] * 100 << std::endl;
        printf("iteration: %d,systemTT: %g, least system TT:%g,gap=%g\n", k, row[0], row[1], row[2]);
    }
    summary << "Step 4: Column Generation for Traffic Assignment" << std::endl;
    summary << "column updating" << std::endl;                    // assignment.h:1057
    for (int n = 0; n < column_updating_iterations; ++n) {        // assignment.h:1060-1075
        double row[4];
        if (dtl_column_update(e, n, row)) fail(dtl_last_error());
        summary << "column updating: iteration= " << n << ", avg travel time = " << row[0] << "(min), optimization obj = " << row[1]
                << ",Relative_gap=" << row[2] << " %" << std::endl; // assignment.h:909-911
    }
    double sys = 0;
    if (dtl_finalize(e, column_generation_iterations, &sys)) fail(dtl_last_error()); // main_api.cpp:748-759
    summary << "Step 4.3: OD estimation " << std::endl;              // main_api.cpp:761-769 (mode dta)
    if (ODME_iterations >= 1) {                                       // Assignment::Demand_ODME, ODME.h:91-516
        const int AT = g->prob.n_agent_types;
        std::vector<float> apce(AT), aocc(AT);
        for (int at = 0; at < AT; ++at) { apce[at] = (float)g->agents[at].pce; aocc[at] = (float)g->agents[at].occ; }
        if (dtl_odme_configure(e, apce.data(), aocc.data(), (int)g->ms_link.size(), g->ms_tau.data(), g->ms_link.data(), g->ms_count.data(),
                               g->ms_ub.data()))
            fail(dtl_last_error());
        summary << "ODME stage" << std::endl;
        {   // sensor_count of the reference only counts records that overwrite an earlier one (ODME.h:158-164)
            int sensor_count = 0;
            std::vector<double> seen((size_t)g->prob.n_periods * g->prob.n_links, 0.0);
            for (size_t i = 0; i < g->ms_link.size(); ++i) {
                double& o = seen[(size_t)g->ms_tau[i] * g->prob.n_links + g->ms_link[i]];
                if (o >= 1) { if (g->ms_ub[i] == 0) { o = g->ms_count[i]; ++sensor_count; } }
                else o = g->ms_count[i];
            }
            summary << "ODME stage: # of sensors =," << sensor_count << std::endl;
        }
        auto odme_row = [&](int s2, const double* r) {                // assignment.h:539-546
            // (the two percentages are float products in the reference)
            summary << "ODME #" << s2 << ", link MAE= " << r[2] << ",link_MAPE: " << (float)r[0] * 100 << "%,system_MPE: " << (float)r[1] * 100
                    << "%,avg_tt = " << r[3] << "(min) " << ",UE gap =" << r[4] << "(min)" << " = (" << r[5] << " %)" << std::endl;
        };
        double prev_gap = 9999999;
        for (int s2 = 0; s2 < ODME_iterations; ++s2) {                // ODME.h:240-260
            double row[8];
            if (dtl_odme_scatter(e, s2, row)) fail(dtl_last_error());
            odme_row(s2, row);
            const double gap_increase = row[0] - prev_gap;
            if (s2 >= 5 && gap_increase > 0.01) {
                summary << "ODME stage terminates with gap increase = " << gap_increase * 100 << "% at iteration = " << s2 << std::endl;
                break;
            }
            prev_gap = row[0];
            if (dtl_odme_update(e)) fail(dtl_last_error());
        }
        double row[8];
        if (dtl_odme_scatter(e, ODME_iterations, row)) fail(dtl_last_error()); // ODME.h:510-512
        odme_row(ODME_iterations, row);
        g->odme_ran = true;
    } else {
        summary << "ODME stage" << std::endl;
    }
    summary << ",# of ODME_iterations=," << ODME_iterations << std::endl;
    if (sa_stage || g->odme_ran) {
        const int L = g->prob.n_links, T = g->prob.n_periods;
        // the "before" record of every link (g_record_corridor_performance_summary(assignment, 0), main_api.cpp:776)
        g->before_volume.assign((size_t)T * L, 0.0); g->before_speed.assign((size_t)T * L, 0.0);
        g->before_doc.assign((size_t)T * L, 0.0); g->before_P.assign((size_t)T * L, 0.0);
        std::vector<double> tt(L), pce(L), doc(L), P(L);
        for (int tau = 0; tau < T; ++tau) {
            if (dtl_get_link_state(e, tau, tt.data(), pce.data(), nullptr, nullptr, doc.data(), nullptr, P.data(), nullptr)) fail(dtl_last_error());
            for (int i = 0; i < L; ++i) {
                const size_t k = (size_t)tau * L + i;
                float spd = g->free_speed[i];                          // output.h:67-69
                if (tt[i] > 0.001f) spd = g->free_speed[i] / std::max(0.000001, tt[i]) * g->fftt[k];
                const float vehicle_volume = pce[i] + g->preload[k];
                g->before_volume[k] = vehicle_volume; g->before_speed[k] = spd; g->before_doc[k] = doc[i]; g->before_P[k] = P[i];
            }
        }
    }
    if (sa_stage) {
        const int L = g->prob.n_links, AT = g->prob.n_agent_types;
        std::vector<int> rt(AT);
        for (int at = 0; at < AT; ++at) rt[at] = g->agents[at].rt_info;
        if (dtl_sa_configure(e, rt.data(), (int)g->sa_link.size(), g->sa_tau.data(), g->sa_link.data(), g->sa_lanes.data())) fail(dtl_last_error());
        if (!g->dms_link.empty() &&                                     // scenario_API.h:165-188
            dtl_sa_set_dms(e, (int)g->dms_link.size(), g->dms_tau.data(), g->dms_link.data(), g->dms_zone.data()))
            fail(dtl_last_error());
        if (dtl_sa_begin(e, nullptr)) fail(dtl_last_error());          // main_api.cpp:779-794, 857
        for (size_t i = 0; i < g->sa_link.size(); ++i) {               // the same lane changes on the host image the writers print
            bool later = false;
            for (size_t j = i + 1; j < g->sa_link.size(); ++j) later = later || (g->sa_tau[j] == g->sa_tau[i] && g->sa_link[j] == g->sa_link[i]);
            if (later) continue;
            const size_t k = (size_t)g->sa_tau[i] * L + g->sa_link[i];
            const float old_lanes = (float)g->nlanes[k];
            g->nlanes[k] = (double)(float)g->sa_lanes[i];
            g->cap[k] *= g->nlanes[k] / std::max(0.00001, (double)old_lanes);
            if (g->nlanes[k] < 0.01) g->fftt[k] = 1440;
        }
        for (int it = 0; it <= S; ++it) {                              // main_api.cpp:864-1034
            double row[2];
            if (dtl_sa_iteration(e, it, row)) fail(dtl_last_error());
            if (it == 0) summary << "Sensitivity analysis stage" << std::endl << ",Iteration,Avg Travel Time(min)" << std::endl;
            summary << it << "," << row[1] << "," << std::endl;
        }
        if (!g->dms_link.empty() && dtl_sa_route_modification(e, nullptr)) fail(dtl_last_error()); // main_api.cpp:1036
        summary << "column updating" << std::endl;                    // main_api.cpp:1040, assignment.h:1057
        for (int n = 0; n < S; ++n) {
            double row[4];
            if (dtl_sa_column_update(e, n, row)) fail(dtl_last_error());
            summary << "column updating: iteration= " << n << ", avg travel time = " << row[0] << "(min), optimization obj = " << row[1]
                    << ",Relative_gap=" << row[2] << " %" << std::endl;
        }
        g->sa_stage = true;
    }
    const double s_loop = since(t_loop);
    const auto t_write = clk::now();
    if (dtl_gmns_write_outputs(g, e, ".")) fail(g_io_err.empty() ? dtl_last_error() : g_io_err);
    const double s_write = since(t_write);
    summary.close();
    printf("wall clock: read input files %.3f s, device image %.3f s, %d + %d assignment iterations %.3f s, write output files %.3f s\n",
           s_read, s_create, column_generation_iterations, column_updating_iterations, s_loop, s_write);
    dtl_destroy(e);
    dtl_gmns_free(g);
    printf("done.\n");
    return 1;
}

int dtalite_main(void)
{
    int s[6];
    if (dtl_gmns_assignment_settings(".", s)) {
        fprintf(stderr, "%s\nProgram stops.\n", g_io_err.c_str());
        return 2;
    }
    network_assignment(1, s[0], s[1], s[2], s[3], s[4], s[5]);
    return 0;
}

} // extern "C"
