------------------------------ MODULE MiniLock ------------------------------
(* Synthetic spec (not from the reference): exercises constructs the ten Kafka files do not use --
   CASE/OTHER, Cardinality (FiniteSets), a primed variable chosen from a set (x' \in S), a primed
   variable read on a right-hand side, BOOLEAN-valued variables with TRUE/FALSE. *)
EXTENDS Integers, FiniteSets

CONSTANTS Procs, Max

VARIABLES holder, waiting, count, flag

NoOne == "none"

TypeOk ==
    /\ holder \in Procs \union {NoOne}
    /\ waiting \subseteq Procs
    /\ count \in 0 .. Max
    /\ flag \in BOOLEAN

Init ==
    /\ holder = NoOne
    /\ waiting = {}
    /\ count = 0
    /\ flag = FALSE

Request(p) ==
    /\ p \notin waiting
    /\ holder # p
    /\ Cardinality(waiting) < 2
    /\ waiting' = waiting \union {p}
    /\ UNCHANGED <<holder, count, flag>>

Grant ==
    /\ holder = NoOne
    /\ waiting # {}
    /\ holder' \in waiting
    /\ waiting' = waiting \ {holder'}
    /\ count' = CASE count < Max -> count + 1
                  [] count = Max -> 0
                  [] OTHER -> count
    /\ flag' = ~flag

Release ==
    /\ holder # NoOne
    /\ holder' = NoOne
    /\ flag' = (count >= 1)
    /\ UNCHANGED <<waiting, count>>

Next ==
    \/ \E p \in Procs : Request(p)
    \/ Grant
    \/ Release

Bounded == Cardinality(waiting) <= 2
HolderNotWaiting == holder = NoOne \/ holder \notin waiting
=============================================================================
