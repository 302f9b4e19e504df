"""Authors the GMNS fixtures under tests/fixtures/ from the reference's datasets (run in the build container):

    python tests/fixtures/make_fixtures.py

  two_corridor   src/test_case_01_two_corridor, inputs verbatim
  corridor_101   dataset/4_101_corridor, inputs verbatim minus measurement.csv / supply_side_scenario.csv
                 (those switch on ODME and scenario analysis, outside the accelerated path)
  corridor_3     dataset/3_corridor_Network/.../network (BASELINE.json configs[1]): node.csv / link.csv verbatim;
                 input_demand.csv with its header renamed to the column-format names; AUTHORED settings.csv
                 (column demand for two agent types, one 4-hour period, BPR) -- the shipped settings name path /
                 activity demand and a subarea, which the committed reference source cannot run
  corridor_3_2p  the same network with TWO demand periods (AM, PM: the reference's MAX_TIMEPERIODS) and an AUTHORED
                 allowed_uses = "sov" on every 7th link, so per-period arrays and the per-agent-type forward-star
                 filter (routing.h:275-297) are exercised
  edge_cases     AUTHORED from scratch: a 4 x 4 grid with two agent types plus the corner cases of the path -- a zone on an
                 island no other zone can reach (unreachable destinations, label 1e15 / predecessor -9999), a zone whose
                 only way out is closed to one agent type (that agent type's tree stops at the first node), a zone without
                 any demand, parallel links between one node pair, demand rows that the reader must skip (o == d, unknown
                 zone, zero volume) and a volume below the 1e-6 back-trace threshold
  toll_signal    AUTHORED: the edge_cases grid with per-agent-type tolls (VDF_toll{agent}1), link penalties (VDF_penalty1), values of
                 time of 7.5 and 31.25 $/h (toll / VOT * 60 with a float VOT, routing.h:229-231, DTA.h:1157) and signalised links
                 (cycle_length >= 1 with start/end green times and a saturation flow rate: the signal branch of the VDF with
                 its float roundings, VDF.h:204-210)
  asu_qvdf       dataset/2_ASU/network (BASELINE.json configs[2]): link.csv verbatim; node.csv with AUTHORED zone ids
                 on every 27th node (the dataset ships without zones); AUTHORED demand and settings with
                 vdf_type=qvdf on every link type, so the queue VDF branch is exercised
"""
import csv
import os
import shutil

REF = os.environ.get("DTL_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def copy(src_dir, dst, names):
    os.makedirs(dst, exist_ok=True)
    for n in names:
        shutil.copy(os.path.join(REF, src_dir, n), os.path.join(dst, n))


def corridor_3():
    src = "dataset/3_corridor_Network/3-corridor_Network_BaceCase/network"
    dst = os.path.join(HERE, "corridor_3")
    copy(src, dst, ["node.csv", "link.csv"])
    with open(os.path.join(REF, src, "input_demand.csv")) as f:
        rows = list(csv.reader(f))
    with open(os.path.join(dst, "demand.csv"), "w") as f:
        f.write("o_zone_id,d_zone_id,volume\n")
        for r in rows[1:]:
            if r and r[0]:
                f.write(",".join(r[:3]) + "\n")
    open(os.path.join(dst, "settings.csv"), "w").write(
        "[assignment],,assignment_mode,column_generation_iterations,column_updating_iterations,number_of_memory_blocks\n"
        ",,dta,20,20,4\n"
        ",,,,,\n"
        "[agent_type],agent_type,name,,vot,flow_type,pce,person_occupancy\n"
        ",sov,passenger,,10,0,1,1\n"
        ",hov,HOV2,,15,0,1,2\n"
        ",,,,,,,\n"
        "[link_type],link_type,link_type_name,,agent_type_blocklist,type_code,traffic_flow_code\n"
        ",1,motorway,,,f,0\n,3,primary,,,a,0\n,4,residential,,,a,0\n,5,secondary,,,a,0\n,6,tertiary,,,a,0\n,11,unclassified,,,a,0\n"
        ",,,,,,\n"
        "[demand_period],demand_period_id,demand_period,,time_period\n"
        ",1,AM,,0600_1000\n"
        ",,,,\n"
        "[demand_file_list],file_sequence_no,file_name,,format_type,demand_period,agent_type,scale_factor\n"
        ",1,demand.csv,,column,AM,sov,6.4\n"
        ",2,demand.csv,,column,AM,hov,1.6\n"
        ",,,,,,,\n"
        "[output_file_configuration],,path_output,major_path_volume_threshold\n"
        ",,1,0.1\n")


def corridor_3_2p():
    src = os.path.join(HERE, "corridor_3")
    dst = os.path.join(HERE, "corridor_3_2p")
    os.makedirs(dst, exist_ok=True)
    for n in ("node.csv", "demand.csv"):
        shutil.copy(os.path.join(src, n), os.path.join(dst, n))
    with open(os.path.join(src, "link.csv")) as f:
        rows = list(csv.reader(f))
    au, li = rows[0].index("allowed_uses"), rows[0].index("link_id")
    for r in rows[1:]:
        if r and r[li] and int(r[li]) % 7 == 3:
            r[au] = "sov"
    with open(os.path.join(dst, "link.csv"), "w", newline="") as f:
        csv.writer(f, lineterminator="\n").writerows(rows)
    s = open(os.path.join(src, "settings.csv")).read()
    s = s.replace(",,dta,20,20,4\n", ",,dta,12,6,4\n")
    s = s.replace(",1,AM,,0600_1000\n", ",1,AM,,0600_1000\n,2,PM,,1500_1900\n")
    s = s.replace(",2,demand.csv,,column,AM,hov,1.6\n", ",2,demand.csv,,column,AM,hov,1.6\n,3,demand.csv,,column,PM,sov,5.0\n"
                  ",4,demand.csv,,column,PM,hov,2.5\n")
    open(os.path.join(dst, "settings.csv"), "w").write(s)


def edge_cases():
    dst = os.path.join(HERE, "edge_cases")
    os.makedirs(dst, exist_ok=True)
    n = 4
    nid = lambda r, c: r * n + c + 1
    zone_of = {nid(0, 0): 1, nid(0, 3): 2, nid(3, 0): 3, nid(3, 3): 4, 17: 5, nid(1, 1): 6}
    with open(os.path.join(dst, "node.csv"), "w") as f:
        f.write("node_id,name,x_coord,y_coord,node_type,ctrl_type,zone_id,geometry\n")
        for r in range(n):
            for c in range(n):
                i = nid(r, c)
                f.write(f"{i},,{c * 0.01:.4f},{r * 0.01:.4f},,,{zone_of.get(i, '')},\n")
        f.write(f"17,,0.1000,0.1000,,,{zone_of[17]},\n18,,0.1100,0.1000,,,,\n")  # the island: nodes 17 <-> 18 only
    links = []
    for r in range(n):
        for c in range(n):
            if c + 1 < n:
                links += [(nid(r, c), nid(r, c + 1)), (nid(r, c + 1), nid(r, c))]
            if r + 1 < n:
                links += [(nid(r, c), nid(r + 1, c)), (nid(r + 1, c), nid(r, c))]
    links += [(17, 18), (18, 17), (nid(1, 2), nid(2, 2))]  # the last one duplicates an existing node pair (parallel link)
    with open(os.path.join(dst, "link.csv"), "w") as f:
        f.write("link_id,name,from_node_id,to_node_id,facility_type,link_type,dir_flag,length,lanes,free_speed,capacity,allowed_uses\n")
        for k, (a, b) in enumerate(links, 1):
            length = 800 + 37 * ((a * 13 + b * 7) % 23)                    # no two paths tie exactly
            uses = "truck" if a == nid(3, 3) else ""                       # zone 4 can be left by trucks only
            speed = (25, 35, 45)[k % 3]
            f.write(f"{k},,{a},{b},,{1 + k % 2},1,{length},{1 + k % 3},{speed},{900 + 200 * (k % 4)},{uses}\n")
    with open(os.path.join(dst, "demand.csv"), "w") as f:
        f.write("o_zone_id,d_zone_id,volume\n")
        f.write("1,2,900\n1,3,450.5\n1,4,300\n2,1,800\n2,3,120\n3,4,610\n4,1,275\n4,2,33\n"
                "1,5,50\n5,1,40\n"          # to / from the island: never reachable
                "2,2,77\n"                    # o == d: skipped by the reader
                "9,1,10\n1,9,10\n"           # unknown zones: skipped
                "3,1,0\n"                     # zero volume
                "3,2,0.0000005\n")            # below the 1e-6 threshold of the back-trace
    open(os.path.join(dst, "settings.csv"), "w").write(
        "[assignment],,assignment_mode,column_generation_iterations,column_updating_iterations,number_of_memory_blocks\n"
        ",,dta,10,4,4\n"
        ",,,,,\n"
        "[agent_type],agent_type,name,,vot,flow_type,pce,person_occupancy\n"
        ",auto,passenger,,10,0,1,1\n"
        ",truck,truck,,25,0,2.5,1\n"
        ",,,,,,,\n"
        "[link_type],link_type,link_type_name,,agent_type_blocklist,type_code,traffic_flow_code\n"
        ",1,major,,,a,0\n,2,minor,,,a,0\n"
        ",,,,,,\n"
        "[demand_period],demand_period_id,demand_period,,time_period\n"
        ",1,AM,,0700_0900\n"
        ",,,,\n"
        "[demand_file_list],file_sequence_no,file_name,,format_type,demand_period,agent_type,scale_factor\n"
        ",1,demand.csv,,column,AM,auto,6\n"
        ",2,demand.csv,,column,AM,truck,1.2\n"
        ",,,,,,,\n"
        "[output_file_configuration],,path_output,major_path_volume_threshold\n"
        ",,1,0.1\n")


def toll_signal():
    """edge_cases' network and demand + toll / penalty / signal columns in link.csv and other values of time"""
    src = os.path.join(HERE, "edge_cases")
    dst = os.path.join(HERE, "toll_signal")
    os.makedirs(dst, exist_ok=True)
    for n in ("node.csv", "demand.csv"):
        shutil.copy(os.path.join(src, n), os.path.join(dst, n))
    with open(os.path.join(src, "link.csv")) as f:
        rows = list(csv.reader(f))
    hdr = rows[0] + ["VDF_tollauto1", "VDF_tolltruck1", "VDF_penalty1", "cycle_length", "start_green_time", "end_green_time",
                     "saturation_flow_rate"]
    out = [hdr]
    for k, r in enumerate(rows[1:], 1):
        toll_auto = f"{0.25 * (k % 4):.2f}" if k % 3 == 0 else "0"
        toll_truck = f"{0.8 + 0.1 * (k % 5):.2f}" if k % 3 == 0 or k % 7 == 0 else "0"
        penalty = f"{0.35 * (1 + k % 3):.2f}" if k % 5 == 0 else "0"
        if k % 4 == 1:      # signalised: 90 s cycle, green from 5 s to 40-55 s
            sig = ["90", "5", str(40 + 5 * (k % 4)), str(1700 + 50 * (k % 3))]
        elif k % 11 == 0:   # a cycle without green times: the whole cycle counts as green, red time 1 s (input.h:3599-3605)
            sig = ["60", "", "", "1900"]
        else:
            sig = ["", "", "", ""]
        out.append(r + [toll_auto, toll_truck, penalty] + sig)
    with open(os.path.join(dst, "link.csv"), "w") as f:
        for r in out:
            f.write(",".join(r) + "\n")
    s = open(os.path.join(src, "settings.csv")).read()
    s = s.replace(",auto,passenger,,10,0,1,1", ",auto,passenger,,7.5,0,1,1").replace(",truck,truck,,25,0,2.5,1", ",truck,truck,,31.25,0,2.5,1")
    assert "7.5" in s and "31.25" in s
    open(os.path.join(dst, "settings.csv"), "w").write(s)


def sa_corridor():
    """3-corridor network with the agent types of dataset/4_101_corridor (p without information, rt with real-time information,
    dms) and a supply_side_scenario.csv that takes lanes away from two motorway links: the sensitivity-analysis stage
    (main_api.cpp:771-1042)"""
    src = os.path.join(HERE, "corridor_3")
    dst = os.path.join(HERE, "sa_corridor")
    os.makedirs(dst, exist_ok=True)
    for n in ("node.csv", "demand.csv", "link.csv"):
        shutil.copy(os.path.join(src, n), os.path.join(dst, n))
    s = open(os.path.join(src, "settings.csv")).read()
    s = s.replace("[agent_type],agent_type,name,,vot,flow_type,pce,person_occupancy\n,sov,passenger,,10,0,1,1\n,hov,HOV2,,15,0,1,2\n",
                  "[agent_type],agent_type,name,,vot,flow_type,pce,person_occupancy,real_time_info\n,p,passenger,,10,0,1,1,0\n"
                  ",rt,passenger,,10,0,1,1,1\n,dms,passenger,,10,0,1,1,2\n")
    s = s.replace(",1,demand.csv,,column,AM,sov,6.4\n,2,demand.csv,,column,AM,hov,1.6\n",
                  ",1,demand.csv,,column,AM,p,4.8\n,2,demand.csv,,column,AM,rt,2.4\n,3,demand.csv,,column,AM,dms,0.8\n")
    assert ",rt,passenger" in s and "AM,rt,2.4" in s
    open(os.path.join(dst, "settings.csv"), "w").write(s)
    with open(os.path.join(dst, "supply_side_scenario.csv"), "w") as f:
        f.write("scenario_type,from_node_id,to_node_id,lanes,demand_period\nsa,4,5,0.5,AM\nsa,5,6,1,AM\n")


def asu_qvdf():
    src = "dataset/2_ASU/network"
    dst = os.path.join(HERE, "asu_qvdf")
    copy(src, dst, ["link.csv"])
    with open(os.path.join(REF, src, "node.csv")) as f:
        rows = list(csv.reader(f))
    hdr = rows[0]
    zi = hdr.index("zone_id")
    zones = []
    out = [hdr]
    for i, r in enumerate(rows[1:]):
        if not r or not r[hdr.index("node_id")]:
            continue
        r = list(r)
        if i % 27 == 5:
            zones.append(len(zones) + 1)
            r[zi] = str(zones[-1])
        out.append(r)
    with open(os.path.join(dst, "node.csv"), "w", newline="") as f:
        csv.writer(f, lineterminator="\n").writerows(out)
    with open(os.path.join(dst, "demand.csv"), "w") as f:
        f.write("o_zone_id,d_zone_id,volume\n")
        for o in zones:
            for d in zones:
                if o != d:
                    f.write(f"{o},{d},{40 + 37 * ((o * 7 + d * 3) % 11)}.5\n")
    with open(os.path.join(REF, src, "link.csv")) as f:
        lrows = list(csv.reader(f))
    lti = lrows[0].index("link_type")
    types = sorted({r[lti] for r in lrows[1:] if len(r) > lti and r[lti]}, key=int)
    lt = "".join(f",{t},type{t},,,a,0,qvdf\n" for t in types)
    open(os.path.join(dst, "settings.csv"), "w").write(
        "[assignment],,assignment_mode,column_generation_iterations,column_updating_iterations,number_of_memory_blocks\n"
        ",,dta,20,0,4\n"
        ",,,,,\n"
        "[agent_type],agent_type,name,,vot,flow_type,pce,person_occupancy\n"
        ",auto,passenger,,10,0,1,1\n"
        ",,,,,,,\n"
        "[link_type],link_type,link_type_name,,agent_type_blocklist,type_code,traffic_flow_code,vdf_type\n" + lt +
        ",,,,,,,\n"
        "[demand_period],demand_period_id,demand_period,,time_period\n"
        ",1,AM,,0700_0800\n"
        ",,,,\n"
        "[demand_file_list],file_sequence_no,file_name,,format_type,demand_period,agent_type\n"
        ",1,demand.csv,,column,AM,auto\n"
        ",,,,,,\n"
        "[output_file_configuration],,path_output,major_path_volume_threshold\n"
        ",,1,0.1\n")


if __name__ == "__main__":
    copy("src/test_case_01_two_corridor", os.path.join(HERE, "two_corridor"), ["node.csv", "link.csv", "demand.csv", "settings.csv"])
    copy("dataset/4_101_corridor", os.path.join(HERE, "corridor_101"), ["node.csv", "link.csv", "demand.csv", "settings.csv"])
    corridor_3()
    corridor_3_2p()
    edge_cases()
    toll_signal()
    sa_corridor()
    asu_qvdf()
    print("fixtures written under", HERE)
