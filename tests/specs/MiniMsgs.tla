---------------------------- MODULE MiniMsgs ----------------------------
(* Second synthetic Sequences spec: a sequence of RECORDS, indexing with a run-time index, a two-parameter
   RECURSIVE operator, \o with a run-time-length right operand. *)
EXTENDS Integers, Sequences, FiniteSets

CONSTANTS Nodes, MaxLen

VARIABLES inbox, cursor, acked

Msg == [from : Nodes, seq : 0 .. 1]

TypeOk ==
    /\ inbox \in Seq(Msg)
    /\ cursor \in 0 .. MaxLen
    /\ acked \subseteq Nodes

RECURSIVE CountFrom(_, _)
CountFrom(s, n) ==
    IF Len(s) = 0 THEN 0
    ELSE (IF Head(s).from = n THEN 1 ELSE 0) + CountFrom(Tail(s), n)

Init ==
    /\ inbox = << >>
    /\ cursor = 0
    /\ acked = {}

Send(n, b) ==
    /\ Len(inbox) < MaxLen
    /\ CountFrom(inbox, n) < 2
    /\ inbox' = Append(inbox, [from |-> n, seq |-> b])
    /\ UNCHANGED <<cursor, acked>>

Advance ==
    /\ cursor < Len(inbox)
    /\ cursor' = cursor + 1
    /\ acked' = acked \union {inbox[cursor + 1].from}
    /\ UNCHANGED inbox

Drop ==
    /\ cursor > 0
    /\ inbox' = Tail(inbox)
    /\ cursor' = cursor - 1
    /\ UNCHANGED acked

Merge ==
    /\ Len(inbox) >= 2
    /\ inbox[1].from = inbox[2].from
    /\ inbox' = <<[from |-> inbox[1].from, seq |-> 1]>> \o SubSeq(inbox, 3, Len(inbox))
    /\ cursor' = IF cursor > 0 THEN cursor - 1 ELSE 0
    /\ UNCHANGED acked

Next ==
    \/ \E n \in Nodes, b \in 0 .. 1 : Send(n, b)
    \/ Advance
    \/ Drop
    \/ Merge

CursorOk == cursor <= Len(inbox)
AckedOk == \A i \in 1 .. cursor : inbox[i].from \in acked
FewPerNode == \A n \in Nodes : CountFrom(inbox, n) <= 2
=============================================================================
