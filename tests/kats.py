"""Known-answer tests transcribed from the reference's DFACompilerTest.java (line numbers cited) and
readme.md.  Shared by the CPU tests (oracle) and the GPU tests (kernels through the Matcher mirror)."""

DOTALL = 0x20

MANY_STATE = "((123)|(234)|(345)|(456)){1,24}"  # IntegrationTest.java:20

# (regex, flags, [strings that match() fully: matches && containedIn && find == (0, len)],
#                [strings that fail(): !containedIn && !matches],
#                [strings with !matches but containedIn])
MATCH_FAIL = [
    ("a", 0, ["a"], ["b", "AB{"], ["ab", "ba"]),                                            # :43-63
    ("xy", 0, ["xy"], ["z", "XY{"], ["xyz", "zxy", "xzxy"]),                                # :82-99
    ("abc", 0, ["abc"], ["d", "AB{", "abd"], ["abcd", "dabc", "abdabc"]),                   # :102-121
    ("abcd", 0, ["abcd"], ["e", "ABC{", "abc", "abce"], ["abcde", "eabcd"]),                # :124-140
    ("abcdefghi", 0, ["abcdefghi"], ["abcd", "abcdefgh"], ["a0cdefghiabcdefghi"]),          # :143-157
    ("a*", 0, ["", "a", "aa", "aaa", "aaaa"], [], ["ab", "e"]),                             # :160-179
    ("ad*g", 0, ["ag", "adg", "adddg"], ["adeg"], []),                                      # :182-197
    ("[0-9A-Za-z]*", 0, ["AB09", "ABC09az"], [], ["AB{"]),                                  # :200-213
    ("(AB)|(BA)", 0, ["AB", "BA"], ["A", "AA", "B", "BB"], ["ABBA"]),                       # :216-234
    ("(A+)|(B+)", 0, ["A", "B", "AA", "BB"], [""], ["AB"]),                                 # :237-255
    (MANY_STATE, 0, ["456", "456456"], ["", "059{"], []),                         # :258-273 (IntegrationTest.MANY_STATE_REGEX_STRING)
    ("A{1,2}", 0, ["A", "AA"], ["", "B", "BB"], ["BAB"]),                                   # :276-293
    ("(AB){1,2}", 0, ["AB", "ABAB"], ["", "BB", "AA"], ["AAB", "ABABAB"]),                  # :296-313
    ("((AB)|(BA)){1,2}", 0, ["BA", "ABBA", "BAAB", "BABA"], [""], []),                      # :316-331
    ("((AB)|(CD)){1,2}AB", 0, ["ABAB", "ABCDAB", "CDAB", "CDCDAB"], [""], []),              # :334-350
    (r"the\s+\w+", 0, ["the a", "the art", "the   art"], ["the", "the ", "the    ", "theart"], ["the   art ", " the a", "a the u"]),  # :353-371
    (MANY_STATE + "ab", 0, ["123ab", "234234ab"], [""], []),                           # :374-386
    ("[0-9]", 0, ["0"], [""], ["0{", "1{"]),                                                # :389-402
    ("[0-9]+", 0, ["0"], [""], ["059{", "12{"]),                                            # :405-418
    ("[؀-ۿ]", 0, ["؀"], ["AB{"], []),                                        # :421-426
    ("A|BCD|E", 0, ["A", "BCD", "E"], ["F"], []),                                           # :429-443
    ("[A-Za-z]+ab", 0, ["Aab", "aab", "AZDab", "ZDaab", "AaDab"], [], []),                  # :446-466
    ("[A-Za-z]+ing", 0, ["bing", "Bing", "zing", "Zing"], [], []),                          # :469-476
    ("[A-Za-z]+abcdef", 0, ["Aabcdef", "aabcdef", "AZDabcdef", "ZDaabcdef"], [], []),       # :479-497
    ("[A-Z]+abcdef", 0, ["Aabcdef", "AZDabcdef"], [], []),                                  # :500-510
    ("[A-Z]+abcdef[A-Z]+", 0, ["AabcdefZ", "AZDabcdefDZA"], [], []),                        # :513-523
    ("Holmes.{0,25}Watson|Watson.{0,25}Holmes", 0, ["HolmesThenWatson"], [], []),           # :543-548
    ("AB.{0,2}12|AB.{0,2}12", 0, ["AB+12"], [], []),                                        # :551-556
    ("((123)|(234)|(345)|(456)|(567)|(678)|(789)|(0987)|(9876)|(8765)|(7654)|(6543)|(5432)|(4321)|(3210)){1,4}", 0,
     ["1232343450987"], [], []),                                                            # :573-578
    ("((123)|(234)|(345)|(456)|(567)|(678)|(789)|(0987)|(9876)|(8765)|(7654)|(6543)|(5432)|(4321)|(3210)){1,8}", 0,
     ["12323434509871232343450987"], [], []),                                               # :580-584
    ("a.*c", 0, ["abc"], [], []),                                                           # :828-830
    ("a.*c", DOTALL, ["abc", "abc\nc"], [], []),                                            # :837-842
]

# (regex, flags, haystack, from, (matched, start, end)) - first find(from, len) on a fresh Matcher
FIND = [
    ("http://.+", 0, "http://www.google.com", 0, (True, 0, 21)),                             # :525-533, readme.md:37-52
    ("http://.+", 0, "http://Γειά σου.com", 0, (True, 0, 19)),  # :535-540
    ("a*baa", 0, "aaaabaa", 3, (True, 3, 7)),                                                 # :785-794
    ("(a*tgc*|t*acg*)*(cg)(a|t)*", 0, "cgatgccgaa", 6, (True, 6, 10)),                        # :803-813
    ("a.*c", 0, "abc\nc", 0, (True, 0, 3)),                                                   # :831-834
    ("a.*c", DOTALL, "abc\nc", 0, (True, 0, 5)),                                              # :838-842
    ("the [Cc]rown", 0, "the Crown", 0, (True, 0, 9)),                                        # :559-562
    ("[a-q][^u-z]{3}x", 0, "aaaax", 0, (True, 0, 5)),                                         # :617-620
]

# (regex, flags, haystack, [all successive (start, end) of while (m.find())])
FIND_ALL = [
    ("a", 0, "aba", [(0, 1), (2, 3)]),                                                        # :66-78
    ("a", 0, "aa", [(0, 1), (1, 2)]),                                                         # :816-825
    ("[a-zA-Z]+ing", 0, "the most perfect reasoning and observing machine that the world has seen",
     [(17, 26), (31, 40)]),                                                                   # :604-614 (count == 2)
]

# regexes whose iterated find() the reference compares with java.util.regex (Python `re` stands in for the JDK)
JDK_DIFFERENTIAL = [
    (".{0,43}A", "@" * 43 + "A"),                                                             # :575-589
    (".{0,47}BCDFHEIJKLAMG", "@" * 47 + "BCDFHEIJKLAMG"),                                     # :635-660
]
