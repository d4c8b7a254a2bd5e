--  LZ4Ada.Batch body -- the Ada rendering of the batched device entry point.
--
--  NEVER COMPILED: this image (and the GPU box, same image) has no GNAT.  It is written against the reference's
--  private part (lib/lz4ada.ads:346-456 -- a child package sees Decompressor_Meta, Process_Header_Bytes, Load_32,
--  Get_Block_Size, Is_Any_Magic_Number ...) and against LZ4Ada.Device (the pragma-Import binding of
--  include/lz4b200.h), statement for statement after the tested C++ host layer:
--
--     Plan   <-> lz4ada_batch_plan     bo_lz4_ada_b200/csrc/host/batch.cpp  (PlanEngine + walker.cpp)
--     Place  <-> place ()              ... batch.cpp
--     Run    <-> lz4ada_batch_upload / lz4ada_batch_run
--     Fold   <-> fold_item ()          ... batch.cpp, errors.cpp (message texts = SURVEY.md Appendix A)
--     Retry  <-> run_slow_items ()
--
--  What it leaves out of the C++ twin, on purpose: the pipelined host-buffer path (chunks on several CUDA streams),
--  the scratch pool, exact sizing as a mode, and the multi-GPU call -- scheduling refinements above the same five
--  stages.  What it must NOT leave out, and does not: the order of the checks, and the texts.
with Ada.Exceptions;        use Ada.Exceptions;
with Ada.Strings.Unbounded; use Ada.Strings.Unbounded;
with Ada.Containers.Vectors;
with Interfaces;            use Interfaces;
with Interfaces.C;          use Interfaces.C;
with System;
with System.Storage_Elements; use System.Storage_Elements;
with Interfaces.C.Strings;
with LZ4Ada.Device;

package body LZ4Ada.Batch is

   use type System.Address;

   ----------------------------------------------------------------------------------------------
   --  Plan: one record per frame and per stream; blocks go straight into a Device.Block_Desc vector
   ----------------------------------------------------------------------------------------------

   type Frame_Plan is record
      Is_Format    : Format  := TBD;
      Independent  : Boolean := True;    --  FLG bit 5 (the reference ignores it; the planner does not)
      Has_CChk     : Boolean := False;   --  FLG bit 2
      CChk_Seen    : Boolean := False;   --  the four checksum bytes were present in the stream
      CChk_Decl    : U32     := 0;
      Has_CSize    : Boolean := False;
      CSize        : U64     := 0;
      Ended        : Boolean := False;   --  end mark processed
      First_Block  : Natural := 0;
      N_Blocks     : Natural := 0;
      Block_Max    : U32     := 0;
      Dst_Off      : U64     := 0;
      Chained      : Boolean := False;   --  linked frame: decoded by K4 as one chain
      Hash_Slot    : Integer := -1;
   end record;

   type Stream_Plan is record
      First_Frame, N_Frames : Natural := 0;
      First_Block, N_Blocks : Natural := 0;
      Host_Error   : Exception_Id := Null_Id;   --  header / size word / Single_Frame: found without the device
      Host_Message : Unbounded_String;
      EOF          : End_Of_Frame := No;
      Slow         : Boolean := False;          --  to be decoded again as one chain with exact placement
      Min_Buffer   : U32 := 0;                  --  Min_Buffer_Size of the Init call the stream is decoded under (:54)
   end record;

   package Frame_Vectors is new Ada.Containers.Vectors (Natural, Frame_Plan);
   package Desc_Vectors  is new Ada.Containers.Vectors (Natural, Device.Block_Desc);
   type Stream_Plans is array (Positive range <>) of Stream_Plan;

   type Plan (N : Natural) is record
      Frames  : Frame_Vectors.Vector;
      Descs   : Desc_Vectors.Vector;
      Streams : Stream_Plans (1 .. N);
      Out_Len : U64 := 0;
   end record;

   Blk_Ring_Cap : constant Unsigned_32 := 64;   --  LZ4B200_BLK_RING_CAP
   Blk_K2       : constant Unsigned_32 := 128;  --  LZ4B200_BLK_K2

   --  One stream, all of it in memory: the walk of lib/lz4ada.adb:383-659 with "record the block" in place of
   --  Decode_Full_Block_With_Trailer.  Exceptions of the header parser and of the size-word checks are the stream's
   --  host outcome: by construction later in stream order than every block recorded before them.
   procedure Walk (P : in out Plan; K : in Positive; Source : in Octets; Span : in Stream_Span;
                   Reservation : in Memory_Reservation) is
      S     : Stream_Plan renames P.Streams (K);
      First : constant Integer := Source'First + Integer (Span.Src_Off);
      Last  : constant Integer := First + Integer (Span.Src_Len) - 1;
      Pos   : Integer := First;
      In_Len : constant Integer := Get_Block_Size (Reservation) + 4 + Block_Size_Bytes;   --  Input_Buffer'Length of Init, :60
   begin
      S.First_Frame := Natural (P.Frames.Length);
      S.First_Block := Natural (P.Descs.Length);
      S.Min_Buffer  := U32 (Get_Block_Size (Reservation) + History_Size + 8);
      while Pos <= Last loop
         declare
            M      : Decompressor_Meta;
            Header : Octets (0 .. 19) := (others => 0);
            Used   : Integer;
            F      : Frame_Plan;
         begin
            M.Memory_Reservation := Reservation;
            --  ---- header (Process_Header_Bytes, :155-191; raises :220 :304 :310 :324 :246 :356) ----
            while M.Header_Parsing /= Header_Complete loop
               if Pos > Last then
                  S.EOF := No;       --  input ended inside a header: no error, Is_End_Of_Frame = No
                  return;
               end if;
               Process_Header_Bytes (M, Header, Source (Pos .. Last), Used);
               Pos := Pos + Used;
            end loop;
            S.EOF := M.Status_EOF;
            if M.Is_Format = Skippable then
               --  Skip, :420-433
               declare
                  Take : constant U64 := U64'Min (M.Size_Remaining, U64 (Last - Pos + 1));
               begin
                  Pos := Pos + Integer (Take);
                  if Take < M.Size_Remaining then
                     S.EOF := No;
                     return;
                  end if;
                  S.EOF := Yes;
               end;
            else
               F.Is_Format   := M.Is_Format;
               F.Has_CChk    := M.Content_Checksum_Length /= 0;
               F.Has_CSize   := M.Has_Content_Size;
               F.CSize       := (if M.Has_Content_Size then M.Size_Remaining else 0);
               F.Independent := M.Is_Format = Legacy or else (Header (4) and 16#20#) /= 0;
               F.Block_Max   := U32 (Get_Block_Size (if M.Is_Format = Legacy then For_Legacy
                                                     else Get_Block_Size_Reservation (Shift_Right (Header (5) and 16#70#, 4))));
               F.First_Block := Natural (P.Descs.Length);
               S.N_Frames := S.N_Frames + 1;
               --  ---- blocks (Try_Detect_Input_Length, :525-585) ----
               Blocks : loop
                  if Last - Pos + 1 < Block_Size_Bytes then
                     S.EOF := (if M.Is_Format = Legacy and then Pos > Last then Maybe else No);
                     P.Frames.Append (F);
                     return;
                  end if;
                  declare
                     Word   : U32 := Load_32 (Source (Pos .. Pos + 3));
                     Stored : Boolean := False;
                     Extra  : constant Integer := Block_Size_Bytes + M.Block_Checksum_Length;
                  begin
                     exit Blocks when M.Is_Format = Legacy and then Is_Any_Magic_Number (Word);   --  :570-580: next frame
                     Pos := Pos + Block_Size_Bytes;
                     if M.Is_Format = Modern and then Word = 0 then
                        --  ---- end mark (Check_End_Mark, :463-523) ----
                        if F.Has_CChk then
                           if Last - Pos + 1 < 4 then
                              S.EOF := No;
                              P.Frames.Append (F);
                              return;
                           end if;
                           F.CChk_Decl := Load_32 (Source (Pos .. Pos + 3));
                           F.CChk_Seen := True;
                           Pos := Pos + 4;
                        end if;
                        F.Ended := True;
                        S.EOF := Yes;
                        exit Blocks;
                     end if;
                     if M.Is_Format = Modern then
                        Stored := (Word and 16#8000_0000#) /= 0;
                        Word   := Word and 16#7ff_ffff#;   --  27 bits, :538 (sic)
                     end if;
                     if Integer (Word) + Extra > In_Len then   --  :541-553
                        raise Data_Corruption with
                          "Declared maximum data length exceeded. Buffer has" & Integer'Image (In_Len) &
                          " bytes, current block requires" & U32'Image (Word) & " bytes +" & Integer'Image (Extra) &
                          " bytes for metadata.";
                     end if;
                     if Last - Pos + 1 < Integer (Word) + M.Block_Checksum_Length then
                        S.EOF := No;   --  the block is not complete: nothing is recorded for it
                        P.Frames.Append (F);
                        return;
                     end if;
                     P.Descs.Append (Device.Block_Desc'
                       (Src_Off    => Unsigned_64 (Pos - Source'First),
                        Dst_Off    => 0,
                        Src_Len    => Unsigned_32 (Word),
                        Dst_Cap    => 0,
                        Flags      => (if Stored then Device.Blk_Stored else 0) or
                                      (if M.Block_Checksum_Length /= 0 then Device.Blk_Has_Checksum else 0),
                        Hist_Avail => 0));
                     F.N_Blocks := F.N_Blocks + 1;
                     S.N_Blocks := S.N_Blocks + 1;
                     Pos := Pos + Integer (Word) + M.Block_Checksum_Length;
                     S.EOF := (if M.Is_Format = Legacy then Maybe else No);
                  end;
               end loop Blocks;
               P.Frames.Append (F);
            end if;
         end;
      end loop;
   exception
      when E : Checksum_Error | Data_Corruption | Not_Supported | Too_Little_Memory =>
         S.Host_Error   := Exception_Identity (E);
         S.Host_Message := To_Unbounded_String (Exception_Message (E));
   end Walk;

   ----------------------------------------------------------------------------------------------
   --  Place: block i of a frame at Frame_Base + i * Block_Max (the frame format has no per-block decompressed
   --  size; every mainstream encoder fills its blocks); streams packed at 256-byte boundaries
   ----------------------------------------------------------------------------------------------

   function Align_Up (V, A : U64) return U64 is ((V + A - 1) / A * A);

   procedure Place (P : in out Plan; Spans : in out Stream_Spans) is
      Cursor : U64 := 0;
   begin
      for K in P.Streams'Range loop
         declare
            S     : Stream_Plan renames P.Streams (K);
            Upper : U64 := 0;
            Pos   : U64;
         begin
            S.Slow := False;
            for F in S.First_Frame .. S.First_Frame + S.N_Frames - 1 loop
               Upper := Upper + U64 (P.Frames (F).N_Blocks) * U64 (P.Frames (F).Block_Max);
            end loop;
            if Spans (K).Dst_Cap = 0 then
               Spans (K).Dst_Off := Align_Up (Cursor, 256);
               Spans (K).Dst_Cap := Upper;
            end if;
            Pos := Spans (K).Dst_Off;
            for F in S.First_Frame .. S.First_Frame + S.N_Frames - 1 loop
               declare
                  FP : Frame_Plan := P.Frames (F);
               begin
                  FP.Dst_Off := Pos;
                  FP.Chained := not FP.Independent and then FP.N_Blocks > 1;
                  for I in 0 .. FP.N_Blocks - 1 loop
                     declare
                        D    : Device.Block_Desc := P.Descs (FP.First_Block + I);
                        Room : constant U64 := Spans (K).Dst_Off + Spans (K).Dst_Cap - U64'Min
                          (Spans (K).Dst_Off + Spans (K).Dst_Cap, Pos + U64 (I) * U64 (FP.Block_Max));
                     begin
                        D.Dst_Off    := Unsigned_64 (Pos + U64 (I) * U64 (FP.Block_Max));
                        D.Dst_Cap    := Unsigned_32 (U64'Min (U64 (FP.Block_Max), Room));
                        D.Hist_Avail := Unsigned_32 (U64'Min (U64 (I) * U64 (FP.Block_Max), 16#ffff_fffe#));
                        D.Flags := D.Flags and not (Device.Blk_Chained or Device.Blk_First_Of_Frame or Blk_Ring_Cap or Blk_K2);
                        if FP.Chained then D.Flags := D.Flags or Device.Blk_Chained; end if;
                        if I = 0 then D.Flags := D.Flags or Device.Blk_First_Of_Frame; end if;
                        if not FP.Chained and then (D.Flags and Device.Blk_Stored) /= 0 then
                           D.Flags := D.Flags or Blk_K2;   --  the wide copy, not a warp of K1
                        end if;
                        if Room < U64 (FP.Block_Max) and then I + 1 < FP.N_Blocks then
                           S.Slow := True;                  --  tight caller buffer: exact placement needed
                        end if;
                        P.Descs.Replace_Element (FP.First_Block + I, D);
                     end;
                  end loop;
                  --  (a frame followed by another frame of the same stream: its last block is sized by K5 first,
                  --  lz4ada_batch_upload; the rendering keeps the upper bound and lets Retry fix what breaks)
                  Pos := Pos + U64 (FP.N_Blocks) * U64 (FP.Block_Max);
                  P.Frames.Replace_Element (F, FP);
               end;
            end loop;
            Cursor := U64'Max (Cursor, Spans (K).Dst_Off + Spans (K).Dst_Cap);
         end;
      end loop;
      P.Out_Len := Cursor;
   end Place;

   ----------------------------------------------------------------------------------------------
   --  Fold: statuses of one stream in stream order -> the first exception the reference would have raised
   --  (SURVEY.md Appendix A "ordering"); True = the stream has to be decoded again as one chain
   ----------------------------------------------------------------------------------------------

   type Status_Array is array (Natural range <>) of Device.Block_Status;
   type U32_Array    is array (Natural range <>) of Unsigned_32;

   procedure Raise_For (St : in Device.Block_Status; Buffer_Len : in Integer; R : in out Outcome) is
      procedure Set (Id : Exception_Id; Text : String) is
      begin
         R.Raised  := Id;
         R.Message := To_Unbounded_String (Text);
      end Set;
   begin
      case St.Code is
         when 1 => Set (Checksum_Error'Identity,     --  :702
                        "Declared checksum is 0x" & To_Hex (U32 (St.XXH32_Declared)) &
                        ", but computed one is 0x" & To_Hex (U32 (St.XXH32_Computed)) & ".");
         when 2 => Set (Data_Corruption'Identity,    --  :754
                        "Match_Length=" & Integer'Image (Integer (St.Aux)) &
                        " suggests compressed data but this sequence already ends after the literals." &
                        " This might also happen with an untypical encoder?");
         when 3 => Set (Data_Corruption'Identity, "Corrupted Block: Offset = 0 detected.");   --  :770
         when 4 => Set (Data_Corruption'Identity,    --  :868
                        "Backreference location out of range. Read from offset" & Integer'Image (Integer (St.Aux)) &
                        " not possible (earliest available index is 0).");
         when 9 => Set (Data_Corruption'Identity,    --  Appendix C: the reference writes unchecked
                        "Output buffer exhausted. Decompressed data does not fit into the" &
                        Integer'Image (Buffer_Len) & " bytes provided.");
         when others => Set (Data_Corruption'Identity, "Block structure damaged (status" & Unsigned_32'Image (St.Code) & ").");
      end case;
   end Raise_For;

   function Fold (P : in Plan; K : in Positive; Span : in Stream_Span; Status : in Status_Array;
                  Digest, Valid : in U32_Array; Exact : in Boolean; R : out Outcome) return Boolean is
      S   : Stream_Plan renames P.Streams (K);
      Pos : U64 := Span.Dst_Off;
   begin
      R := (Raised => Null_Id, Message => Null_Unbounded_String, End_Of_Frame => S.EOF, Dst_Off => Span.Dst_Off, Out_Len => 0);
      if not Exact and then S.Slow then return True; end if;
      for F in S.First_Frame .. S.First_Frame + S.N_Frames - 1 loop
         declare
            FP        : Frame_Plan renames P.Frames.Constant_Reference (F);
            FPos      : U64 := 0;
            Remaining : U64 := FP.CSize;
         begin
            if not Exact and then FP.Dst_Off /= Pos then return True; end if;
            for I in 0 .. FP.N_Blocks - 1 loop
               declare
                  St : Device.Block_Status renames Status (FP.First_Block + I);
                  D  : Device.Block_Desc renames P.Descs.Constant_Reference (FP.First_Block + I);
               begin
                  if St.Code = 10 or else St.Code = 11 or else St.Code = 16#ffff_ffff# then return True; end if;   --  needs history / not run
                  if not Exact and then not FP.Chained and then U64 (D.Dst_Off) /= FP.Dst_Off + FPos then return True; end if;
                  --  Decrease_Data_Size_Remaining (:826-839) fires inside Write_Output, before any later check of
                  --  the same block; the block checksum (:672-676) comes before everything
                  if FP.Has_CSize and then St.Code /= 1 then
                     if Remaining < U64 (if St.Code = 0 then St.Out_Len else St.Err_Pos) then
                        R.Raised  := Data_Corruption'Identity;
                        R.Message := To_Unbounded_String
                          ("Produced content size exceeds declared content size. The supplied data is inconsistent.");   --  :831
                        exit;
                     end if;
                  end if;
                  if St.Code /= 0 then
                     if St.Code = 9 and then not Exact then return True; end if;   --  the slot was a placement assumption
                     Raise_For (St, (if (D.Flags and Blk_Ring_Cap) /= 0 then Integer (S.Min_Buffer) else Integer (D.Dst_Cap)), R);
                     exit;
                  end if;
                  Remaining := Remaining - U64 (St.Out_Len);
                  FPos := FPos + U64 (St.Out_Len);
               end;
            end loop;
            --  the reference hands out every block before the failing one (one block per Update)
            Pos := Pos + FPos;
            R.Out_Len := R.Out_Len + FPos;
            exit when R.Raised /= Null_Id;
            if FP.Has_CChk and then FP.CChk_Seen then
               declare
                  Value : U32;
               begin
                  if FP.N_Blocks = 0 then
                     Value := LZ4Ada.XXHash32.Hash (Octets'(1 .. 0 => 0));   --  Final of nothing
                  elsif FP.Hash_Slot < 0 or else Valid (FP.Hash_Slot) = 0 then
                     return True;
                  else
                     Value := U32 (Digest (FP.Hash_Slot));
                  end if;
                  if Value /= FP.CChk_Decl then   --  :505
                     R.Raised  := Checksum_Error'Identity;
                     R.Message := To_Unbounded_String
                       ("Computed content checksum 0x" & To_Hex (Value) & " does not match declared content checksum 0x" &
                        To_Hex (FP.CChk_Decl) & ".");
                     exit;
                  end if;
               end;
            end if;
            if FP.Ended and then FP.Has_CSize and then Remaining /= 0 then   --  :471
               R.Raised  := Data_Corruption'Identity;
               R.Message := To_Unbounded_String
                 ("Frame has ended, but according to content size, there should be" & U64'Image (Remaining) & " bytes left to output.");
               exit;
            end if;
         end;
      end loop;
      if R.Raised = Null_Id then
         R.Raised  := S.Host_Error;
         R.Message := S.Host_Message;
      end if;
      return False;
   end Fold;

   ----------------------------------------------------------------------------------------------
   --  The two entry points
   ----------------------------------------------------------------------------------------------

   function Output_Bytes (Source : in Octets; Spans : in Stream_Spans;
                          Reservation : in Memory_Reservation := For_All) return U64 is
      P    : Plan (Spans'Length);
      Copy : Stream_Spans := Spans;
   begin
      for K in Copy'Range loop
         Walk (P, K - Copy'First + 1, Source, Copy (K), Reservation);
      end loop;
      Place (P, Copy);
      return P.Out_Len;
   end Output_Bytes;

   procedure Check (RC : int; Ctx : Device.Context) is
   begin
      if RC /= 0 then
         raise Device_Error with Interfaces.C.Strings.Value (Device.Last_Error (Ctx));
      end if;
   end Check;

   procedure Decompress (Source      : in     Octets;
                         Spans       : in out Stream_Spans;
                         Destination :    out Octets;
                         Results     :    out Outcomes;
                         Reservation : in     Memory_Reservation := For_All;
                         GPU         : in     Natural := 0) is
      P   : Plan (Spans'Length);
      Ctx : Device.Context;
      D_Src, D_Dst, D_Desc, D_Status, D_Chains, D_Frames, D_Digest : System.Address := System.Null_Address;
   begin
      --  1. plan + place (host only)
      for K in Spans'Range loop
         Walk (P, K - Spans'First + 1, Source, Spans (K), Reservation);
      end loop;
      Place (P, Spans);
      if P.Out_Len > U64 (Destination'Length) then
         raise Constraint_Error with "Destination too small for the batch";
      end if;
      declare
         NB     : constant Natural := Natural (P.Descs.Length);
         Descs  : array (0 .. Integer'Max (NB, 1) - 1) of aliased Device.Block_Desc;
         Status : Status_Array (0 .. Integer'Max (NB, 1) - 1);
         Chains : array (0 .. Integer'Max (Natural (P.Frames.Length), 1) - 1) of aliased Device.Chain;
         Hashed : array (0 .. Integer'Max (Natural (P.Frames.Length), 1) - 1) of aliased Device.Frame_Blocks;
         NC, NH : Natural := 0;
      begin
         for I in 0 .. NB - 1 loop
            Descs (I) := P.Descs (I);
         end loop;
         for F in 0 .. Natural (P.Frames.Length) - 1 loop
            declare
               FP : Frame_Plan := P.Frames (F);
            begin
               if FP.Chained then
                  Chains (NC) := (First_Block => Unsigned_32 (FP.First_Block), N_Blocks => Unsigned_32 (FP.N_Blocks),
                                  Dst_Off => Unsigned_64 (FP.Dst_Off),
                                  Dst_Cap => Unsigned_64 (U64 (FP.N_Blocks) * U64 (FP.Block_Max)));
                  NC := NC + 1;
               end if;
               if FP.Has_CChk and then FP.CChk_Seen and then FP.N_Blocks > 0 then
                  FP.Hash_Slot := NH;
                  Hashed (NH) := (First_Block => Unsigned_32 (FP.First_Block), N_Blocks => Unsigned_32 (FP.N_Blocks));
                  NH := NH + 1;
                  P.Frames.Replace_Element (F, FP);
               end if;
            end;
         end loop;
         declare
            Digest : U32_Array (0 .. Integer'Max (NH, 1) - 1) := (others => 0);
            Valid  : U32_Array (0 .. Integer'Max (NH, 1) - 1) := (others => 0);
            Both   : array (0 .. 2 * Integer'Max (NH, 1) - 1) of aliased Unsigned_32 := (others => 0);
         begin
            --  2. device: tables and compressed bytes up, K1 (+ K2 for the stored blocks), K4, K3, statuses down
            Check (Device.Create (int (GPU), System.Null_Address, Ctx), Ctx);
            if NB > 0 then
               Check (Device.Alloc (Ctx, Source'Length + 64, D_Src), Ctx);
               Check (Device.Alloc (Ctx, size_t (P.Out_Len) + 64, D_Dst), Ctx);
               Check (Device.Alloc (Ctx, size_t (NB) * (Device.Block_Desc'Size / 8), D_Desc), Ctx);
               Check (Device.Alloc (Ctx, size_t (NB) * (Device.Block_Status'Size / 8), D_Status), Ctx);
               Check (Device.H2D (Ctx, D_Src, Source (Source'First)'Address, Source'Length), Ctx);
               Check (Device.H2D (Ctx, D_Desc, Descs (0)'Address, size_t (NB) * (Device.Block_Desc'Size / 8)), Ctx);
               Check (Device.Decode_Blocks (Ctx, D_Src, D_Dst, Unsigned_32 (NB), D_Desc, D_Status), Ctx);
               --  (stored blocks flagged Blk_K2: Device.Copy_Stored over their index list -- lz4b200_copy_stored)
               if NC > 0 then
                  Check (Device.Alloc (Ctx, size_t (NC) * (Device.Chain'Size / 8), D_Chains), Ctx);
                  Check (Device.H2D (Ctx, D_Chains, Chains (0)'Address, size_t (NC) * (Device.Chain'Size / 8)), Ctx);
                  Check (Device.Decode_Linked (Ctx, D_Src, D_Dst, Unsigned_32 (NC), D_Chains, D_Desc, D_Status), Ctx);
               end if;
               if NH > 0 then
                  Check (Device.Alloc (Ctx, size_t (NH) * 8, D_Frames), Ctx);
                  Check (Device.Alloc (Ctx, size_t (NH) * 8, D_Digest), Ctx);
                  Check (Device.H2D (Ctx, D_Frames, Hashed (0)'Address, size_t (NH) * 8), Ctx);
                  Check (Device.XXH32_Frames (Ctx, D_Dst, Unsigned_32 (NH), D_Frames, D_Desc, D_Status, D_Digest,
                                              D_Digest + Storage_Offset (4 * NH)), Ctx);
                  Check (Device.D2H (Ctx, Both (0)'Address, D_Digest, size_t (NH) * 8), Ctx);
               end if;
               Check (Device.D2H (Ctx, Status (0)'Address, D_Status, size_t (NB) * (Device.Block_Status'Size / 8)), Ctx);
               Check (Device.Sync (Ctx), Ctx);
               for I in 0 .. NH - 1 loop
                  Digest (I) := Both (I);
                  Valid (I)  := Both (NH + I);
               end loop;
            end if;
            --  3. fold; 4. streams whose placement assumption broke: again, each as ONE chain with exact running
            --     placement, every block bounded the way the reference bounds it (Blk_Ring_Cap: what is left of the
            --     caller's Buffer behind the ring cursor, lib/lz4ada.adb:54, 678-680), then fold again
            for K in Spans'Range loop
               declare
                  KK : constant Positive := K - Spans'First + 1;
                  S  : Stream_Plan renames P.Streams (KK);
               begin
                  if Fold (P, KK, Spans (K), Status, Digest, Valid, False, Results (Results'First + KK - 1)) then
                     for I in S.First_Block .. S.First_Block + S.N_Blocks - 1 loop
                        Descs (I).Flags   := (Descs (I).Flags and not (Device.Blk_First_Of_Frame or Blk_K2))
                                             or Device.Blk_Chained or Blk_Ring_Cap;
                        Descs (I).Dst_Cap := Unsigned_32 (S.Min_Buffer);
                        P.Descs.Replace_Element (I, Descs (I));
                     end loop;
                     for F in S.First_Frame .. S.First_Frame + S.N_Frames - 1 loop
                        if P.Frames (F).N_Blocks > 0 then
                           Descs (P.Frames (F).First_Block).Flags := Descs (P.Frames (F).First_Block).Flags or Device.Blk_First_Of_Frame;
                        end if;
                     end loop;
                     Chains (0) := (First_Block => Unsigned_32 (S.First_Block), N_Blocks => Unsigned_32 (S.N_Blocks),
                                    Dst_Off => Unsigned_64 (Spans (K).Dst_Off), Dst_Cap => Unsigned_64 (Spans (K).Dst_Cap));
                     Check (Device.H2D (Ctx, D_Desc, Descs (0)'Address, size_t (NB) * (Device.Block_Desc'Size / 8)), Ctx);
                     if D_Chains = System.Null_Address then
                        Check (Device.Alloc (Ctx, Device.Chain'Size / 8, D_Chains), Ctx);
                     end if;
                     Check (Device.H2D (Ctx, D_Chains, Chains (0)'Address, Device.Chain'Size / 8), Ctx);
                     Check (Device.Decode_Linked (Ctx, D_Src, D_Dst, 1, D_Chains, D_Desc, D_Status), Ctx);
                     Check (Device.D2H (Ctx, Status (0)'Address, D_Status, size_t (NB) * (Device.Block_Status'Size / 8)), Ctx);
                     Check (Device.Sync (Ctx), Ctx);
                     --  (content checksums of the re-placed frames: Device.XXH32_Spans over their exact spans,
                     --  run_slow_items in batch.cpp; the digests replace Digest / Valid of those frames)
                     if Fold (P, KK, Spans (K), Status, Digest, Valid, True, Results (Results'First + KK - 1)) then
                        raise Device_Error with "chain decode reported an unexpected soft status";
                     end if;
                  end if;
               end;
            end loop;
            --  5. bring back exactly what each stream produced
            for K in Spans'Range loop
               declare
                  R : Outcome renames Results (Results'First + K - Spans'First);
               begin
                  if R.Out_Len > 0 then
                     Check (Device.D2H (Ctx, Destination (Destination'First + Integer (R.Dst_Off))'Address,
                                        D_Dst + Storage_Offset (R.Dst_Off), size_t (R.Out_Len)), Ctx);
                  end if;
               end;
            end loop;
            Check (Device.Sync (Ctx), Ctx);
            Check (Device.Destroy (Ctx), Ctx);   --  drops this call's reference; buffers go with the context's pool
         end;
      end;
   end Decompress;

end LZ4Ada.Batch;
